"""PPO loss, hand-derived gradient, clip_by_global_norm + Adam, epoch/minibatch loops.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates /root/reference/src/lerax/algorithm/ppo.py:396-467 (ppo_loss), :471-492
(train_batch), :494-521 (train_epoch), :523-528 (explained_variance), :530-561 (train),
/root/reference/src/lerax/buffer/base_buffer.py:38-103 (flatten i = n*T + t, batch_indices,
gather) and optax 0.2.6 (third-party, /root/reference/uv.lock:1783-1784):
    chain(clip_by_global_norm(max_norm), inject_hyperparams(adam)(lr)), ppo.py:192-194
    clip:  g <- g if ||g||_2 < max_norm else g / ||g||_2 * max_norm
    adam:  count += 1; m <- b1 m + (1-b1) g; v <- b2 v + (1-b2) g^2
           p <- p - lr * (m / (1-b1^count)) / (sqrt(v / (1-b2^count)) + eps)   (eps_root = 0)
The gradient follows JAX sub-gradient conventions: jnp.minimum / jnp.clip split ties 1/2-1/2.
"parity unpinned" against JAX (not installable here); cross-checked against torch autograd
and torch.optim.Adam in tests/test_oracle_torch_crosscheck.py.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import distributions as D
from . import policy as P

F32_EPS = float(np.finfo(np.float32).eps)  # jnp.finfo(advantages.dtype).eps, ppo.py:426


@dataclass
class Hyper:
    clip_coefficient: float = 0.2
    clip_value_loss: bool = False
    entropy_loss_coefficient: float = 0.0
    value_loss_coefficient: float = 0.5
    max_grad_norm: float = 0.5
    normalize_advantages: bool = True
    learning_rate: float = 3e-4
    b1: float = 0.9
    b2: float = 0.999
    eps: float = 1e-8
    # 0: PPO clipped surrogate (ppo.py:429-434); 1: -mean(log_prob * advantage) of A2C / REINFORCE
    # (algorithm/a2c.py:401, algorithm/reinforce.py:391)
    surrogate: int = 0


class NonFiniteError(FloatingPointError):
    """eqx.error_if(~isfinite) in ppo_loss, ppo.py:413-417."""


def flatten_time_major(x_tn):
    """[T, N, ...] (kernel layout) -> flat [N*T, ...] in reference order i = n*T + t
    (base_buffer.py:38-62 applied to the reference's [N, T, ...] buffer)."""
    return np.swapaxes(x_tn, 0, 1).reshape((-1,) + x_tn.shape[2:])


def _tie_weights(a, b):
    """d min(a,b): (1,0) if a<b, (0,1) if b<a, (.5,.5) on ties (JAX convention)."""
    wa = np.where(a < b, 1.0, np.where(a == b, 0.5, 0.0)).astype(a.dtype)
    return wa, (1 - wa).astype(a.dtype)


def _clip_grad(x, lo, hi):
    """d clip(x, lo, hi)/dx with 1/2 at the bounds (clip = minimum(maximum(x, lo), hi))."""
    g = np.where((x > lo) & (x < hi), 1.0, 0.0)
    g = np.where((x == lo) | (x == hi), 0.5, g)
    return g.astype(x.dtype)


def ppo_loss_and_grad(spec: P.PolicySpec, params, mb, hp: Hyper, eps_dtype_eps=None):
    """mb: dict(obs [B,d], act [B], logp [B], val [B], ret [B], adv [B]).
    Returns (loss, stats[5] = approx_kl,total,policy,value,entropy losses, grads flat)."""
    dtype = params.dtype
    dt = dtype.type
    B = mb["obs"].shape[0]
    invB = dt(1.0) / dt(B)
    cache = []
    value, head = P.forward(spec, params, mb["obs"], cache)
    _, ls = spec.offsets()
    if spec.is_box:
        scale = np.exp(params[ls])
        loc = head[:, 0]
        logp = D.normal_log_prob(loc, scale, mb["act"])
        entropy = np.broadcast_to(D.normal_entropy(scale), logp.shape)
    else:
        lp = D.log_softmax(head)
        logp = np.take_along_axis(lp, mb["act"][:, None].astype(np.int64), axis=1)[:, 0]
        prob = np.exp(lp)
        entropy = -np.sum(prob * lp, axis=-1)
    if not (np.isfinite(value).all() and np.isfinite(logp).all() and np.isfinite(entropy).all()):
        raise NonFiniteError("Non-finite values/log_probs/entropy.")

    log_ratio = logp - mb["logp"]
    ratio = np.exp(log_ratio)
    approx_kl = np.mean(ratio - log_ratio) - dt(1.0)

    adv = mb["adv"]
    if hp.normalize_advantages:
        eps = dt(F32_EPS if eps_dtype_eps is None else eps_dtype_eps)
        adv = (adv - np.mean(adv)) / (np.std(adv) + eps)

    c = dt(hp.clip_coefficient)
    lo, hi = dt(1.0) - c, dt(1.0) + c
    if hp.surrogate == 1:
        policy_loss = -np.mean(logp * adv)
        dlogp = -invB * adv
    else:
        s1 = adv * ratio
        s2 = adv * np.clip(ratio, lo, hi)
        policy_loss = -np.mean(np.minimum(s1, s2))
        w1, w2 = _tie_weights(s1, s2)
        dlogp = -invB * (w1 * adv * ratio + w2 * adv * _clip_grad(ratio, lo, hi) * ratio)

    if hp.clip_value_loss:
        dv_old = value - mb["val"]
        clipped = mb["val"] + np.clip(dv_old, -c, c)
        e1 = value - mb["ret"]
        e2 = clipped - mb["ret"]
        q1, q2 = e1 * e1, e2 * e2
        value_loss = np.mean(np.minimum(q1, q2)) / dt(2.0)
        u1, u2 = _tie_weights(q1, q2)
        dvalue_l = invB * (u1 * e1 + u2 * e2 * _clip_grad(dv_old, -c, c))
    else:
        e1 = value - mb["ret"]
        value_loss = np.mean(e1 * e1) / dt(2.0)
        dvalue_l = invB * e1

    entropy_loss = -np.mean(entropy)
    cv, ch = dt(hp.value_loss_coefficient), dt(hp.entropy_loss_coefficient)
    loss = policy_loss + value_loss * cv + entropy_loss * ch

    dvalue = cv * dvalue_l
    dlog_std = None
    if spec.is_box:
        z = (mb["act"] - loc) / scale
        dhead = (dlogp * z / scale)[:, None]
        # d logp / d log_std = z^2 - 1 ; d entropy / d log_std = 1 per row, mean -> 1
        dlog_std = np.sum(dlogp * (z * z - dt(1.0))) - ch
    else:
        onehot = np.zeros_like(lp)
        np.put_along_axis(onehot, mb["act"][:, None].astype(np.int64), 1.0, axis=1)
        dhead = dlogp[:, None] * (onehot - prob)
        dH = -prob * (lp + entropy[:, None])  # d entropy / d logits
        dhead = dhead + (-ch * invB) * dH
    grads = P.backward(spec, params, cache, dvalue, dhead.astype(dtype), dlog_std)
    stats = np.array([approx_kl, loss, policy_loss, value_loss, entropy_loss], dtype=dtype)
    return loss, stats, grads


@dataclass
class AdamState:
    m: np.ndarray
    v: np.ndarray
    count: int = 0

    @classmethod
    def init(cls, params):
        return cls(np.zeros_like(params), np.zeros_like(params), 0)

    def copy(self):
        return AdamState(self.m.copy(), self.v.copy(), self.count)


def global_norm(g):
    return np.sqrt(np.sum(g * g))


def clip_by_global_norm(g, max_norm):
    n = global_norm(g)
    dt = g.dtype.type
    if n < dt(max_norm):
        return g, n
    return g / n * dt(max_norm), n


def adam_update(params, g, st: AdamState, hp: Hyper, lr=None):
    dt = params.dtype.type
    b1, b2 = dt(hp.b1), dt(hp.b2)
    lr = hp.learning_rate if lr is None else lr
    if callable(lr):
        lr = lr(st.count)
    count = st.count + 1
    m = b1 * st.m + (dt(1.0) - b1) * g
    v = b2 * st.v + (dt(1.0) - b2) * (g * g)
    m_hat = m / (dt(1.0) - b1 ** dt(count))
    v_hat = v / (dt(1.0) - b2 ** dt(count))
    upd = m_hat / (np.sqrt(v_hat) + dt(hp.eps))
    new_params = params + dt(-1.0) * dt(lr) * upd
    return new_params.astype(params.dtype), AdamState(m, v, count)


def train_batch(spec, params, st, mb, hp):
    """ppo.py:471-492."""
    _, stats, g = ppo_loss_and_grad(spec, params, mb, hp)
    g, gnorm = clip_by_global_norm(g, hp.max_grad_norm)
    params, st = adam_update(params, g, st, hp)
    return params, st, stats, gnorm


def gather(flat, idx):
    return {k: v[idx] for k, v in flat.items()}


def train_epoch(spec, params, st, flat, idx, hp):
    """ppo.py:494-521. idx [num_batches, B] of flat indices i = n*T + t."""
    stats = []
    for row in idx:
        params, st, s, _ = train_batch(spec, params, st, gather(flat, row), hp)
        stats.append(s)
    return params, st, np.mean(np.stack(stats), axis=0, dtype=params.dtype)


def explained_variance(returns, values):
    """ppo.py:523-528 over the whole buffer."""
    dt = returns.dtype.type
    var = np.var(returns)
    return dt(1.0) - np.var(returns - values) / (var + dt(F32_EPS))


def train(spec, params, st, flat, idx_epochs, hp):
    """ppo.py:530-561. idx_epochs [E, num_batches, B].  Returns params, state, log dict."""
    per_epoch = []
    for idx in idx_epochs:
        params, st, s = train_epoch(spec, params, st, flat, idx, hp)
        per_epoch.append(s)
    s = np.mean(np.stack(per_epoch), axis=0, dtype=params.dtype)
    log = dict(
        approx_kl=s[0], loss=s[1], policy_loss=s[2], value_loss=s[3], entropy_loss=s[4],
        explained_variance=explained_variance(flat["ret"], flat["val"]),
    )
    return params, st, log


def batch_indices(total, batch_size, rng: np.random.Generator | None):
    """base_buffer.py:64-88 with a NumPy permutation in place of jr.permutation."""
    idx = np.arange(total) if rng is None else rng.permutation(total)
    trim = total - (total % batch_size)
    return idx[:trim].reshape(-1, batch_size).astype(np.int32)
