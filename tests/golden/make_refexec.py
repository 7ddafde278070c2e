#!/usr/bin/env python
"""Generate tests/golden/refexec_*.npz by EXECUTING THE REFERENCE'S OWN SOURCE (unmodified, from
/root/reference/src/lerax) over the third-party stand-ins of oracle/_refexec (JAX is not installable here).

What runs is the reference's code: env `initial / dynamics / transition / reward / terminal / observation`,
`TimeLimit`, `PPOStepState.initial`, `PPO.collect_rollout` (-> `PPO.step`, `filter_scan`, `filter_cond`,
`RolloutBuffer.compute_returns_and_advantages` -> `GAE.__call__`) under `eqx.filter_vmap` exactly as `PPO.iteration`
calls it (algorithm/ppo.py:360-368), `PPO.ppo_loss_grad`, `PPO.train` (-> `train_epoch`, `train_batch`,
`flatten_axes`, `batch_indices`, `gather`, optax chain), and `MLPActorCriticPolicy` with injected parameters.
Random draws are injected by key path.  The fixtures pin oracle/ (tests/test_oracle_refexec.py) and, through the C ABI,
the CUDA kernels (tests/test_gpu_refexec.py); they travel to the GPU box, /root/reference does not.

Run (authoring container only):  python tests/golden/make_refexec.py
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import policy as OP  # noqa: E402  (parameter layout + init only)
from oracle._refexec import loader  # noqa: E402

ns = loader.load()
import torch  # noqa: E402

eqx, jr, jnp = ns.eqx, ns.jr, ns.jnp

N_ENVS, T_STEPS, N_BATCHES, N_EPOCHS, GAMMA, LAM = 6, 24, 3, 2, 0.99, 0.95


def np32(x):
    if isinstance(x, torch.Tensor):
        x = x.detach().numpy()
    return np.asarray(x)


# ---------------------------------------------------------------------------------- cases
def make_env(kind, mes):
    env = {"cartpole": ns.CartPole, "pendulum": ns.Pendulum, "acrobot": ns.Acrobot, "mountain_car": ns.MountainCar,
           "continuous_mountain_car": ns.ContinuousMountainCar}[kind]()
    return ns.TimeLimit(env, mes) if mes else env


CASES = [
    # name, env kind, TimeLimit steps, policy kwargs, (obs_dim, n_out, is_box, state_dim)
    ("cartpole", "cartpole", 0, {}, (4, 2, False, 4)),
    ("cartpole_tl9", "cartpole", 9, {}, (4, 2, False, 4)),
    ("pendulum_tl7", "pendulum", 7, {}, (3, 1, True, 2)),
    ("acrobot", "acrobot", 0, {}, (6, 3, False, 4)),
    ("mountain_car_tl11", "mountain_car", 11, {}, (2, 3, False, 2)),
    ("continuous_mountain_car", "continuous_mountain_car", 0, {}, (2, 1, True, 2)),
    # non-default widths / depths (the generic-width kernels): action_depth 1 -> the action MLP is one Linear(16,16)
    ("pendulum_w32", "pendulum", 0, dict(feature_width=32, feature_depth=1, value_width=24, value_depth=3,
                                         action_width=40, action_depth=1, feature_size=8), (3, 1, True, 2)),
]

HYPERS = [
    dict(),
    dict(clip_value_loss=True, entropy_loss_coefficient=0.01, clip_coefficient=0.1),
    dict(normalize_advantages=False, value_loss_coefficient=1.0, max_grad_norm=0.05),
]


def inject_params(policy, flat):
    """Overwrite encoder / value_head / action_head float leaves with `flat` (Equinox leaf order)."""
    off = [0]

    def repl(x):
        if eqx.is_inexact_array(x):
            n = x.numel()
            v = torch.from_numpy(flat[off[0]:off[0] + n].reshape(tuple(x.shape)).copy())
            off[0] += n
            return v
        return x

    nets = (policy.encoder, policy.value_head, policy.action_head)
    new = tuple(eqx.tree_map(repl, m) for m in nets)
    assert off[0] == flat.size, (off[0], flat.size)
    return eqx.tree_at(lambda p: (p.encoder, p.value_head, p.action_head), policy, new)


def flat_params(tree):
    """encoder ++ value_head ++ action_head float leaves of a policy-shaped pytree (policy or gradients)."""
    out = []
    for m in (tree.encoder, tree.value_head, tree.action_head):
        out += [np32(x).reshape(-1) for x in eqx.tree_leaves(m) if eqx.is_inexact_array(x)]
    return np.concatenate(out).astype(np.float32)


class Noise:
    """Draws for one case; the handler maps a key path to its slice (see oracle/_refexec/shim/jax/random.py)."""

    def __init__(self, rng, n_out, is_box, k):
        self.init_u = rng.random((N_ENVS, k)).astype(np.float32)
        self.reset_u = rng.random((T_STEPS, N_ENVS, k)).astype(np.float32)
        self.action = (rng.standard_normal((T_STEPS, N_ENVS)) if is_box
                       else rng.gumbel(size=(T_STEPS, N_ENVS, n_out))).astype(np.float32)
        total = N_ENVS * T_STEPS
        self.perm = np.stack([rng.permutation(total) for _ in range(N_EPOCHS)]).astype(np.int32)
        self.k = k

    def __call__(self, path, kind, shape):
        # ('reset', 0, n, 0)                      PPO.reset -> split(step_key, N)[n] -> PPOStepState.initial env_key
        # ('iter', 0, n, 0, t, 0 | 6)             iteration rollout_key -> env n -> collect key -> step t -> action | env_reset
        # ('iter', 1, e)                          train_key -> epoch e -> batch_indices permutation
        if path[0] == "reset" and kind == "uniform":
            return self.init_u[path[2]][: int(np.prod(shape))].reshape(shape)
        if path[0] == "iter" and path[1] == 0 and len(path) == 6:
            _, _, n, _, t, j = path
            if j == 0 and kind in ("gumbel", "normal"):
                return self.action[t, n].reshape(shape)
            if j == 6 and kind == "uniform":
                return self.reset_u[t, n][: int(np.prod(shape))].reshape(shape)
        if path[0] == "iter" and path[1] == 1 and kind == "permutation":
            return self.perm[path[2]]
        if path[0] == "probe":
            return self.probe[kind][: int(np.prod(shape))].reshape(shape)
        raise KeyError(f"no injected noise for key path {path} ({kind}, {shape})")


def unwrap_state(s):
    """(y [k], t, step_count) of a (possibly TimeLimit-wrapped) env state."""
    sc = getattr(s, "step_count", None)
    inner = s.env_state if sc is not None else s
    return inner.y, inner.t, (sc if sc is not None else torch.zeros(inner.t.shape, dtype=torch.int32))


def env_probe(env, kind, rng, k, is_box, n_act, M=48):
    """Single transitions from scattered states (some beyond the termination thresholds)."""
    inner = env.env if hasattr(env, "env") else env
    scale = {"cartpole": [3.0, 3.0, 0.3, 3.0], "pendulum": [4.0, 9.0], "acrobot": [3.3, 3.3, 14.0, 12.0],
             "mountain_car": [1.3, 0.08], "continuous_mountain_car": [1.3, 0.08]}[kind]
    y = (rng.uniform(-1, 1, (M, k)) * np.asarray(scale)).astype(np.float32)
    if kind in ("mountain_car", "continuous_mountain_car"):
        y[:, 0] = np.clip(y[:, 0] - 0.3, -1.2, 0.6)
    act = (rng.uniform(-2.5, 2.5, M).astype(np.float32) if is_box else rng.integers(0, n_act, M).astype(np.int32))
    t0 = rng.uniform(0, 3, M).astype(np.float32)
    key = jr.make_key(("unused",))
    out = {k2: [] for k2 in ("dyn", "y1", "t1", "rew", "term", "obs", "clip")}
    y_wide = (y * 3.0).astype(np.float32)  # env.clip alone, far outside the valid ranges (angle wrap, velocity / position limits)
    state0 = inner.initial(key=jr.make_key(("probe",)))
    for i in range(M):
        s = eqx.tree_at(lambda q: (q.y, q.t), state0, (torch.from_numpy(y[i].copy()), torch.tensor(t0[i])))
        a = torch.tensor(act[i])
        out["dyn"].append(np32(inner.dynamics(s.t, s.y, a)))
        s1 = inner.transition(s, a, key=key)
        out["y1"].append(np32(s1.y))
        out["t1"].append(np32(s1.t))
        out["rew"].append(np32(inner.reward(s, a, s1, key=key)))
        out["term"].append(np32(inner.terminal(s1, key=key)))
        out["obs"].append(np32(inner.observation(s, key=key)))
        out["clip"].append(np32(inner.clip(torch.from_numpy(y_wide[i].copy()))))
    res = {f"probe_{k2}": np.stack(v) for k2, v in out.items()}
    res.update(probe_y=y, probe_t=t0, probe_act=act, probe_y_wide=y_wide)
    return res


def run_case(name, kind, mes, pol_kw, shape_info, seed):
    obs_dim, n_out, is_box, k = shape_info
    rng = np.random.default_rng(seed)
    env = make_env(kind, mes)
    spec = OP.make_spec(obs_dim, n_out, is_box, **pol_kw)
    params = OP.init_params(spec, seed=seed, log_std_init=-0.3 if is_box else 0.0)
    noise = Noise(rng, n_out, is_box, k)
    noise.probe = {"uniform": rng.random(k).astype(np.float32)}
    jr.reset_registry()
    jr.set_noise_handler(noise)
    policy = inject_params(ns.MLPActorCriticPolicy(env, key=jr.make_key(("init",)), **pol_kw), params)
    assert np.array_equal(flat_params(policy), params)
    fx = dict(params=params, init_u=noise.init_u, reset_u=noise.reset_u, action_noise=noise.action, perm=noise.perm,
              gamma=np.float32(GAMMA), lam=np.float32(LAM), max_episode_steps=np.int32(mes))
    fx.update(env_probe(env, kind, rng, k, is_box, n_out))
    fx["probe_initial"] = np32(unwrap_state(env.initial(key=jr.make_key(("probe",))))[0])
    fx["probe_initial_u"] = noise.probe["uniform"]

    # ---- PPO.reset + the rollout half of PPO.iteration (algorithm/ppo.py:318-344, 353-368), unmodified
    algo = ns.PPO(num_envs=N_ENVS, num_steps=T_STEPS, num_epochs=N_EPOCHS, num_batches=N_BATCHES, gae_lambda=LAM,
                  gamma=GAMMA)
    callback = ns.CallbackList(callbacks=[])
    state = algo.reset(env, policy, key=jr.make_key(("reset",)), callback=callback)
    y0, t0, sc0 = unwrap_state(state.step_state.env_state)
    fx.update(y0=np32(y0).T.copy(), t0=np32(t0), sc0=np32(sc0).astype(np.int32))  # y [k, N]
    rollout_key, train_key, _ = jr.split(jr.make_key(("iter",)), 3)
    step_state, buf = eqx.filter_vmap(algo.collect_rollout, in_axes=(None, None, eqx.if_array(0), None, 0))(
        state.env, state.policy, state.step_state, callback, jr.split(rollout_key, N_ENVS))
    y1, t1, sc1 = unwrap_state(step_state.env_state)
    tm = lambda x: np.swapaxes(np32(x), 0, 1).copy()  # noqa: E731  reference [N, T, ...] -> kernel layout [T, N, ...]
    fx.update(obs=tm(buf.observations), act=tm(buf.actions).astype(np.float32 if is_box else np.int32), rew=tm(buf.rewards),
              done=tm(buf.dones), logp=tm(buf.log_probs), val=tm(buf.values), adv=tm(buf.advantages), ret=tm(buf.returns),
              y_end=np32(y1).T.copy(), t_end=np32(t1), sc_end=np32(sc1).astype(np.int32))
    # last_value as collect_rollout computes it (ppo.py:311-312) for the final carried states
    lv = [np32(policy.value(None, env.observation(eqx.tree_map(lambda x: x[n] if eqx.is_array(x) else x, step_state.env_state),
                                                  key=jr.make_key(("unused",))))[1]) for n in range(N_ENVS)]
    fx["last_value"] = np.stack(lv).astype(np.float32)

    # ---- ppo_loss value / stats / gradient of single minibatches (algorithm/ppo.py:396-469)
    flat = buf.flatten_axes()
    B = algo.batch_size
    for h, hk in enumerate(HYPERS):
        a2 = ns.PPO(num_envs=N_ENVS, num_steps=T_STEPS, num_epochs=N_EPOCHS, num_batches=N_BATCHES, gae_lambda=LAM,
                    gamma=GAMMA, **hk)
        idx = torch.from_numpy(noise.perm[0][h * B:(h + 1) * B].copy())
        (loss, stats), grads = a2.ppo_loss_grad(policy, flat.gather(idx), a2.normalize_advantages, a2.clip_coefficient,
                                                a2.clip_value_loss, a2.value_loss_coefficient, a2.entropy_loss_coefficient)
        fx[f"loss_idx_{h}"] = noise.perm[0][h * B:(h + 1) * B]
        fx[f"loss_stats_{h}"] = np.array([np32(stats.approx_kl), np32(stats.total_loss), np32(stats.policy_loss),
                                          np32(stats.value_loss), np32(stats.entropy_loss)], np.float32)
        fx[f"loss_grad_{h}"] = flat_params(grads)
        assert abs(float(loss) - float(stats.total_loss)) == 0.0

    # ---- PPO.train: epochs x minibatches of train_batch with the optax chain (algorithm/ppo.py:471-561)
    for h, hk in enumerate(HYPERS[:2]):
        a2 = ns.PPO(num_envs=N_ENVS, num_steps=T_STEPS, num_epochs=N_EPOCHS, num_batches=N_BATCHES, gae_lambda=LAM,
                    gamma=GAMMA, learning_rate=1e-3, **hk)
        opt_state = a2.optimizer.init(eqx.filter(policy, eqx.is_inexact_array))
        new_policy, new_opt, log = a2.train(policy, opt_state, buf, key=train_key)
        adam = new_opt[1].inner_state[0]
        fx[f"train_params_{h}"] = flat_params(new_policy)
        if h == 0:
            fx[f"train_mu_{h}"], fx[f"train_nu_{h}"] = flat_params(adam.mu), flat_params(adam.nu)
        fx[f"train_count_{h}"] = np.int32(int(adam.count))
        fx[f"train_log_{h}"] = np.array([np32(log[k2]) for k2 in ("approx_kl", "loss", "policy_loss", "value_loss",
                                                                  "entropy_loss", "explained_variance")], np.float32)
    fx["hyper_sets"] = np.array(repr(HYPERS))
    fx["policy_kwargs"] = np.array(repr(pol_kw))
    return fx


def gae_cases(rng):
    """GAE.__call__ (advantage/gae.py:36-66) on random strips, beyond the reference's own known answers."""
    out = {}
    for i, (T, gamma, lam, p_done) in enumerate([(17, 0.99, 0.95, 0.1), (64, 0.9, 0.8, 0.3), (5, 1.0, 1.0, 0.0),
                                                 (33, 0.97, 0.0, 0.2)]):
        r, v = rng.standard_normal(T).astype(np.float32), rng.standard_normal(T).astype(np.float32)
        d = rng.random(T) < p_done
        lv = np.float32(rng.standard_normal())
        adv, ret = ns.GAE(gamma=gamma, lam=lam)(torch.from_numpy(r), torch.from_numpy(v), torch.from_numpy(d),
                                                torch.tensor(lv))
        out.update({f"gae{i}_rew": r, f"gae{i}_val": v, f"gae{i}_done": d, f"gae{i}_last": lv,
                    f"gae{i}_gl": np.array([gamma, lam], np.float32), f"gae{i}_adv": np32(adv), f"gae{i}_ret": np32(ret)})
    out["n"] = np.int32(4)
    return out


def main():
    for i, (name, kind, mes, pol_kw, shape_info) in enumerate(CASES):
        fx = run_case(name, kind, mes, pol_kw, shape_info, seed=100 + i)
        path = os.path.join(HERE, f"refexec_{name}.npz")
        np.savez_compressed(path, **fx)
        print(f"{name}: dones {int(fx['done'].sum())}/{fx['done'].size}, grad norm {np.linalg.norm(fx['loss_grad_0']):.4f},"
              f" {os.path.getsize(path) / 1024:.0f} KB")
    np.savez_compressed(os.path.join(HERE, "refexec_gae.npz"), **gae_cases(np.random.default_rng(7)))


if __name__ == "__main__":
    main()
