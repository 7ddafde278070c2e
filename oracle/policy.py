"""MLPActorCriticPolicy: parameter layout, init, forward and hand-derived backward.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates /root/reference/src/lerax/policy/actor_critic/mlp.py:64-178 and
/root/reference/src/lerax/policy/actor.py:40-88,194-261 on top of equinox 0.13.6's
``eqx.nn.MLP`` / ``eqx.nn.Linear`` (third-party, /root/reference/uv.lock:420-422):
``Linear``: y = W[out,in] @ x + b, W,b ~ U(+-1/sqrt(in)); ``MLP(depth)``: depth hidden
Linears each followed by the activation, then a final Linear with identity output.

    encoder     : obs_dim -> feature_width x feature_depth -> feature_size
    value_head  : feature_size -> value_width x value_depth -> scalar
    action_head : MLP(feature_size -> action_width x max(0, action_depth-1) -> feature_size)
                  then Linear(feature_size -> n_out), NO activation in between (actor.py:248-261)
    log_std     : scalar, Box (Pendulum) only (actor.py:60-63)

Flat parameter vector = Equinox leaf order (utils.py:309-329): for each chain
(encoder, value_head, action_head.mlp ++ action_dist.mapping), for each Linear:
weight[out,in] row-major, then bias[out]; finally log_std.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

ENC, VAL, ACT = 0, 1, 2


@dataclass(frozen=True)
class PolicySpec:
    obs_dim: int
    n_out: int  # number of logits (Discrete) or 1 (scalar Box)
    is_box: bool
    dims: tuple  # (enc_dims, val_dims, act_dims); each a tuple of layer widths incl. input
    relu: tuple  # per chain, per Linear: activation after it?

    @property
    def n_params(self) -> int:
        n = 0
        for chain in self.dims:
            for i in range(len(chain) - 1):
                n += chain[i] * chain[i + 1] + chain[i + 1]
        return n + (1 if self.is_box else 0)

    def offsets(self):
        """[(chain, layer, w_off, b_off, in, out)] in flat order, and log_std offset (or -1)."""
        out = []
        off = 0
        for c, chain in enumerate(self.dims):
            for i in range(len(chain) - 1):
                k, n = chain[i], chain[i + 1]
                out.append((c, i, off, off + k * n, k, n))
                off += k * n + n
        return out, (off if self.is_box else -1)


def make_spec(obs_dim, n_out, is_box, *, feature_size=16, feature_width=64, feature_depth=2,
              value_width=64, value_depth=2, action_width=64, action_depth=2) -> PolicySpec:
    """mlp.py:64-110 + actor.py:232-246 (depth = max(0, action_depth - 1))."""
    enc = (obs_dim,) + (feature_width,) * feature_depth + (feature_size,)
    val = (feature_size,) + (value_width,) * value_depth + (1,)
    ad = max(0, action_depth - 1)
    act = (feature_size,) + (action_width,) * ad + (feature_size,) + (n_out,)
    relu = (
        (True,) * feature_depth + (False,),
        (True,) * value_depth + (False,),
        (True,) * ad + (False, False),
    )
    return PolicySpec(obs_dim, n_out, is_box, (enc, val, act), relu)


def init_params(spec: PolicySpec, seed=0, log_std_init=0.0, dtype=np.float32):
    """eqx.nn.Linear init: U(-1/sqrt(in), 1/sqrt(in)) for weight and bias."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p = np.zeros(spec.n_params, dtype=dtype)
    offs, ls = spec.offsets()
    for (_, _, w, b, k, n) in offs:
        lim = 1.0 / np.sqrt(k)
        p[w:w + k * n] = rng.uniform(-lim, lim, size=k * n)
        p[b:b + n] = rng.uniform(-lim, lim, size=n)
    if ls >= 0:
        p[ls] = log_std_init
    return p


def _chain_layers(spec, params, c):
    offs, _ = spec.offsets()
    out = []
    for (cc, i, w, b, k, n) in offs:
        if cc == c:
            out.append((params[w:w + k * n].reshape(n, k), params[b:b + n], spec.relu[c][i]))
    return out


def _run_chain(spec, params, c, x, cache=None):
    """x: [B, in].  cache (list) receives (input, pre-activation-output) per layer."""
    for (W, b, relu) in _chain_layers(spec, params, c):
        z = x @ W.T + b
        if cache is not None:
            cache.append((x, z))
        x = np.maximum(z, 0) if relu else z
    return x


def forward(spec: PolicySpec, params, obs, cache=None):
    """obs [B, obs_dim] -> (value [B], head [B, n_out]) ; head = logits or loc."""
    c_enc = [] if cache is not None else None
    c_val = [] if cache is not None else None
    c_act = [] if cache is not None else None
    feat = _run_chain(spec, params, ENC, obs, c_enc)
    value = _run_chain(spec, params, VAL, feat, c_val)[:, 0]
    head = _run_chain(spec, params, ACT, feat, c_act)
    if cache is not None:
        cache.extend([c_enc, c_val, c_act])
    return value, head


def value_only(spec, params, obs):
    feat = _run_chain(spec, params, ENC, obs)
    return _run_chain(spec, params, VAL, feat)[:, 0]


def _backprop_chain(spec, params, c, cache, dout, grads):
    """dout [B, out_last] -> d input [B, in_first]; accumulates into flat grads."""
    offs, _ = spec.offsets()
    mine = [(w, b, k, n, i) for (cc, i, w, b, k, n) in offs if cc == c]
    for (w, b, k, n, i), (x, z) in zip(reversed(mine), reversed(cache)):
        if spec.relu[c][i]:
            dout = dout * (z > 0)
        W = params[w:w + k * n].reshape(n, k)
        grads[w:w + k * n] += (dout.T @ x).reshape(-1)
        grads[b:b + n] += dout.sum(axis=0)
        dout = dout @ W
    return dout


def backward(spec: PolicySpec, params, cache, dvalue, dhead, dlog_std=None):
    """Given dL/dvalue [B], dL/dhead [B, n_out] (and dL/dlog_std scalar) return flat grads."""
    grads = np.zeros_like(params)
    c_enc, c_val, c_act = cache
    dfeat = _backprop_chain(spec, params, VAL, c_val, dvalue[:, None], grads)
    dfeat = dfeat + _backprop_chain(spec, params, ACT, c_act, dhead, grads)
    _backprop_chain(spec, params, ENC, c_enc, dfeat, grads)
    _, ls = spec.offsets()
    if ls >= 0 and dlog_std is not None:
        grads[ls] += dlog_std
    return grads
