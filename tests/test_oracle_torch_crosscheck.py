"""Cross-check the unpinned parts of the oracle (policy forward, ppo_loss gradient,
clip_by_global_norm, Adam, rollout bookkeeping) against independent implementations
available in this image: torch autograd (fp64), torch.optim.Adam, torch clip_grad_norm_."""

import numpy as np
import pytest
import torch

from oracle import envs as E
from oracle import gae as G
from oracle import policy as P
from oracle import ppo as O
from oracle import rollout as R


def spec_for(kind, **kw):
    if kind == "cartpole":
        return E.cartpole(), P.make_spec(4, 2, False, **kw)
    if kind == "pendulum":
        return E.pendulum(), P.make_spec(3, 1, True, **kw)
    return E.acrobot(), P.make_spec(6, 3, False, **kw)


def test_param_counts():
    # CartPole 12 995, Acrobot 13 140, Pendulum 12 915 (+log_std), 256-wide 149 811
    assert spec_for("cartpole")[1].n_params == 12995
    assert spec_for("acrobot")[1].n_params == 13140
    assert spec_for("pendulum")[1].n_params == 12915
    assert spec_for("pendulum", feature_width=256, value_width=256, action_width=256)[1].n_params == 149811


def torch_forward(spec, p, obs):
    offs, ls = spec.offsets()

    def chain(c, x):
        for (cc, i, w, b, k, n) in offs:
            if cc == c:
                x = x @ p[w:w + k * n].reshape(n, k).T + p[b:b + n]
                if spec.relu[c][i]:
                    x = torch.relu(x)
        return x

    f = chain(0, obs)
    return chain(1, f)[:, 0], chain(2, f)


def torch_loss(spec, p, mb, hp):
    value, head = torch_forward(spec, p, mb["obs"])
    _, ls = spec.offsets()
    if spec.is_box:
        dist = torch.distributions.Normal(head[:, 0], torch.exp(p[ls]))
        logp = dist.log_prob(mb["act"])
        ent = dist.entropy()
    else:
        dist = torch.distributions.Categorical(logits=head)
        logp = dist.log_prob(mb["act"])
        ent = dist.entropy()
    lr = logp - mb["logp"]
    ratio = torch.exp(lr)
    kl = (ratio - lr).mean() - 1
    adv = mb["adv"]
    if hp.normalize_advantages:
        adv = (adv - adv.mean()) / (adv.std(unbiased=False) + O.F32_EPS)
    c = hp.clip_coefficient
    if hp.surrogate == 1:
        pl = -(logp * adv).mean()  # a2c.py:401, reinforce.py:391
    else:
        pl = -torch.minimum(adv * ratio, adv * torch.clamp(ratio, 1 - c, 1 + c)).mean()
    if hp.clip_value_loss:
        cl = mb["val"] + torch.clamp(value - mb["val"], -c, c)
        vl = torch.minimum((value - mb["ret"]) ** 2, (cl - mb["ret"]) ** 2).mean() / 2
    else:
        vl = ((value - mb["ret"]) ** 2).mean() / 2
    el = -ent.mean()
    loss = pl + hp.value_loss_coefficient * vl + hp.entropy_loss_coefficient * el
    return loss, torch.stack([kl, loss, pl, vl, el])


def make_minibatch(kind, B, dtype, seed=0, **kw):
    env, spec = spec_for(kind, **kw)
    rng = np.random.default_rng(seed)
    params = P.init_params(spec, seed=seed, log_std_init=-0.3, dtype=dtype)
    obs = rng.standard_normal((B, spec.obs_dim)).astype(dtype)
    if spec.is_box:
        act = rng.uniform(-2, 2, B).astype(dtype)
    else:
        act = rng.integers(0, spec.n_out, B).astype(np.int32)
    mb = dict(obs=obs, act=act, logp=(rng.standard_normal(B) * 0.3 - 1).astype(dtype),
              val=rng.standard_normal(B).astype(dtype), ret=rng.standard_normal(B).astype(dtype),
              adv=rng.standard_normal(B).astype(dtype))
    return spec, params, mb


@pytest.mark.parametrize("kind", ["cartpole", "pendulum", "acrobot"])
@pytest.mark.parametrize("clip_v,ch,sur", [(False, 0.0, 0), (True, 0.01, 0), (False, 0.01, 1)])
def test_loss_and_grad_match_torch_autograd_fp64(kind, clip_v, ch, sur):
    spec, params, mb = make_minibatch(kind, 96, np.float64, seed=3)
    hp = O.Hyper(clip_value_loss=clip_v, entropy_loss_coefficient=ch, surrogate=sur)
    loss, stats, g = O.ppo_loss_and_grad(spec, params, mb, hp)
    p = torch.tensor(params, requires_grad=True)
    tmb = {k: torch.tensor(v) for k, v in mb.items()}
    if not spec.is_box:
        tmb["act"] = tmb["act"].long()
    tl, ts = torch_loss(spec, p, tmb, hp)
    tl.backward()
    np.testing.assert_allclose(stats, ts.detach().numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(g, p.grad.numpy(), rtol=1e-8, atol=1e-12)
    assert np.abs(g).max() > 0


def test_fp32_oracle_close_to_fp64_twin():
    spec, p64, mb64 = make_minibatch("cartpole", 256, np.float64, seed=5)
    p32 = p64.astype(np.float32)
    mb32 = {k: (v.astype(np.float32) if v.dtype == np.float64 else v) for k, v in mb64.items()}
    hp = O.Hyper()
    _, s64, g64 = O.ppo_loss_and_grad(spec, p64, mb64, hp)
    _, s32, g32 = O.ppo_loss_and_grad(spec, p32, mb32, hp)
    np.testing.assert_allclose(s32, s64, rtol=2e-5, atol=2e-6)
    assert np.abs(g32 - g64).max() <= 1e-5 * np.abs(g64).max() + 1e-7


def test_clip_and_adam_match_torch():
    spec, params, mb = make_minibatch("cartpole", 64, np.float64, seed=7)
    hp = O.Hyper(max_grad_norm=0.05)  # force the clip branch
    st = O.AdamState.init(params)
    p_t = torch.tensor(params, requires_grad=True)
    opt = torch.optim.Adam([p_t], lr=hp.learning_rate, betas=(hp.b1, hp.b2), eps=hp.eps)
    tmb = {k: torch.tensor(v) for k, v in mb.items()}
    tmb["act"] = tmb["act"].long()
    p_o = params.copy()
    for _ in range(5):
        p_o, st, _, gnorm = O.train_batch(spec, p_o, st, mb, hp)
        opt.zero_grad()
        torch_loss(spec, p_t, tmb, hp)[0].backward()
        tn = torch.nn.utils.clip_grad_norm_([p_t], hp.max_grad_norm)
        # torch scales by max_norm/(norm+1e-6); optax by max_norm/norm: compensate exactly
        p_t.grad.mul_((float(tn) + 1e-6) / float(tn)) if float(tn) > hp.max_grad_norm else None
        np.testing.assert_allclose(gnorm, float(tn), rtol=1e-10)
        opt.step()
    np.testing.assert_allclose(p_o, p_t.detach().numpy(), rtol=1e-9, atol=1e-12)
    assert st.count == 5


def test_no_clip_below_threshold():
    g = np.array([0.3, 0.0, -0.39], np.float32)
    out, n = O.clip_by_global_norm(g, 0.5)
    assert n < 0.5 and np.array_equal(out, g)
    out, n = O.clip_by_global_norm(g * 10, 0.5)
    np.testing.assert_allclose(np.sqrt((out**2).sum()), 0.5, rtol=1e-6)


def test_nonfinite_raises():
    spec, params, mb = make_minibatch("cartpole", 8, np.float32)
    mb["obs"][0, 0] = np.inf
    with pytest.raises(O.NonFiniteError):
        O.ppo_loss_and_grad(spec, params, mb, O.Hyper())


def test_flatten_order_and_batch_indices():
    T, N = 3, 2
    x = np.arange(T * N).reshape(T, N)  # x[t, n] = t*N + n
    flat = O.flatten_time_major(x)
    for n in range(N):
        for t in range(T):
            assert flat[n * T + t] == x[t, n]  # base_buffer.py:38-62: i = n*T + t
    idx = O.batch_indices(10, 4, np.random.default_rng(0))
    assert idx.shape == (2, 4) and len(set(idx.ravel().tolist())) == 8
    assert O.batch_indices(10, 4, None).tolist() == [[0, 1, 2, 3], [4, 5, 6, 7]]


@pytest.mark.parametrize("kind", ["cartpole", "pendulum", "acrobot"])
def test_rollout_bookkeeping(kind):
    """Step-order semantics of ppo.py:199-288 on a short rollout with TimeLimit."""
    env, spec = spec_for(kind)
    env.max_episode_steps = 5
    N, T = 6, 17
    rng = np.random.default_rng(11)
    params = P.init_params(spec, seed=1)
    y = E.initial_from_uniform(env, rng.random((env.state_dim, N)).astype(np.float32))
    noise = (rng.standard_normal((T, N)) if spec.is_box else rng.gumbel(size=(T, N, spec.n_out))).astype(np.float32)
    ru = rng.random((T, N, env.state_dim)).astype(np.float32)
    out = R.rollout(env, spec, params, y, np.zeros(N, np.float32), np.zeros(N, np.int32), T, 0.99, noise, ru,
                    trace=True)
    assert out["done"].any()
    for s in range(T):
        # obs is the observation of the PRE-transition state
        np.testing.assert_array_equal(out["obs"][s], E.observation(env, out["y_trace"][s]).T)
        v, head = P.forward(spec, params, out["obs"][s])
        np.testing.assert_allclose(out["val"][s], v, rtol=1e-6)
        # reward stored is bootstrapped only on truncation, from the pre-reset next state
        trunc = out["done"][s] & ~E.terminal(env, out["y_next_trace"][s])
        r = E.reward(env, out["y_trace"][s], out["act"][s], out["y_next_trace"][s])
        vb = P.value_only(spec, params, E.observation(env, out["y_next_trace"][s]).T.copy())
        np.testing.assert_allclose(out["rew"][s], np.where(trunc, r + np.float32(0.99) * vb, r), rtol=1e-6)
        if s + 1 < T:
            nxt = out["y_trace"][s + 1]
            y0 = E.initial_from_uniform(env, ru[s].T)
            np.testing.assert_array_equal(nxt, np.where(out["done"][s][None], y0, out["y_next_trace"][s]))
            assert (out["t_trace"][s + 1][out["done"][s]] == 0).all()
    if spec.is_box:
        assert (np.abs(out["act"]) <= 2.0).all()
    adv, ret = G.gae(out["rew"], out["val"], out["done"], out["last_value"], 0.99, 0.95)
    assert np.isfinite(adv).all() and np.allclose(ret, adv + out["val"])


def test_oracle_ppo_learns_cartpole():
    """Same statement as tests/algorithm/test_ppo.py:36-69: trained >= untrained return."""
    env, spec = spec_for("cartpole")
    N, T, E_, NB = 4, 128, 4, 8
    rng = np.random.default_rng(7)
    params = P.init_params(spec, seed=7)

    def evaluate(p):
        tot = []
        for ep in range(8):
            y = E.initial_from_uniform(env, np.random.default_rng(100 + ep).random((4, 1)).astype(np.float32))
            t = np.zeros(1, np.float32)
            ret = 0
            for _ in range(500):
                _, head = P.forward(spec, p, E.observation(env, y).T.copy())
                y, t = E.transition(env, y, t, np.argmax(head, -1))
                ret += 1
                if E.terminal(env, y)[0]:
                    break
            tot.append(ret)
        return np.mean(tot)

    before = evaluate(params)
    st = O.AdamState.init(params)
    hp = O.Hyper()
    y = E.initial_from_uniform(env, rng.random((4, N)).astype(np.float32))
    t = np.zeros(N, np.float32)
    sc = np.zeros(N, np.int32)
    for it in range(4096 // (N * T)):
        out = R.rollout(env, spec, params, y, t, sc, T, 0.99,
                        rng.gumbel(size=(T, N, 2)).astype(np.float32), rng.random((T, N, 4)).astype(np.float32))
        y, t, sc = out["y"], out["t"], out["step_count"]
        adv, ret = G.gae(out["rew"], out["val"], out["done"], out["last_value"], 0.99, 0.95)
        flat = {k: O.flatten_time_major(v) for k, v in
                dict(obs=out["obs"], act=out["act"], logp=out["logp"], val=out["val"], ret=ret, adv=adv).items()}
        idx = np.stack([O.batch_indices(N * T, N * T // NB, rng) for _ in range(E_)])
        params, st, log = O.train(spec, params, st, flat, idx, hp)
        assert all(np.isfinite(v) for v in log.values())
    after = evaluate(params)
    assert np.isfinite(before) and np.isfinite(after) and after >= before
