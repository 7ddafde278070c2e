"""PPO rollout collection (PPO.step scanned over T, vmapped over N) with injected noise.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates /root/reference/src/lerax/algorithm/ppo.py:199-288 (step), :290-316
(collect_rollout) and :360-368 (vmap over envs).  Order of events per env-step
(SURVEY.md Appendix A.1):
    obs   = env.observation(y)                               # pre-transition state
    a_raw, value, logp = policy.action_and_value(obs, noise) # logp of the UNCLIPPED sample
    a     = clip(a_raw, low, high) if Box else a_raw         # ppo.py:228-235
    y', t'= env.transition(y, t, a)                          # one Tsit5 step + env.clip
    r     = env.reward(y, a, y')
    done  = env.terminal(y') | env.truncate(state')          # TimeLimit: step_count+1 >= max
    r_b   = r + gamma * V(obs(y')) if truncate else r        # ppo.py:248-259, BEFORE reset
    state = env.initial(reset noise) if done else state'     # ppo.py:261-263, t=0, step_count=0
    emit (obs, a, r_b, done, logp, value)                    # ppo.py:276-288
After T steps: last_value = V(obs(final state)) (ppo.py:311-312); GAE follows (oracle.gae).
Arrays are time-major [T, N] (the CUDA layout); the reference's logical layout is the
transpose [N, T] (vmap axis outermost).
"""

from __future__ import annotations

import numpy as np

from . import distributions as D
from . import envs as E
from . import policy as P


def act(env_spec, pol_spec, params, obs_bn, noise):
    """policy.action_and_value (mlp.py:133-152).  obs_bn [N, d]; noise [N, n_out] Gumbel or [N] eps.
    Returns (env_action, stored_action, value, logp)."""
    value, head = P.forward(pol_spec, params, obs_bn)
    dt = obs_bn.dtype.type
    if pol_spec.is_box:
        _, ls = pol_spec.offsets()
        scale = np.exp(params[ls])
        loc = head[:, 0]
        a_raw = D.normal_sample(loc, scale, noise)
        logp = D.normal_log_prob(loc, scale, a_raw)
        lo, hi = E.action_bounds(env_spec, dt)
        a = np.clip(a_raw, lo, hi)
        return a, a, value, logp
    lp = D.log_softmax(head)
    a = D.categorical_sample(lp, noise)
    logp = np.take_along_axis(lp, a[:, None].astype(np.int64), axis=1)[:, 0]
    return a, a, value, logp


def rollout(env_spec, pol_spec, params, y, t, step_count, T, gamma, noise_action, reset_uniform,
            trace=False):
    """Collect T steps for N envs.

    y [k,N], t [N], step_count [N] int32 : carried env state (modified copies returned)
    noise_action [T,N,n_out] (Gumbel) or [T,N] (normal eps); reset_uniform [T,N,k] in [0,1)
    """
    dtype = y.dtype
    dt = dtype.type
    N = y.shape[1]
    d = env_spec.obs_dim
    y, t, step_count = y.copy(), t.copy(), step_count.copy()
    out = dict(
        obs=np.zeros((T, N, d), dtype),
        act=np.zeros((T, N), dtype if env_spec.is_box else np.int32),
        rew=np.zeros((T, N), dtype),
        done=np.zeros((T, N), bool),
        logp=np.zeros((T, N), dtype),
        val=np.zeros((T, N), dtype),
    )
    if trace:
        out["y_trace"] = np.zeros((T,) + y.shape, dtype)
        out["t_trace"] = np.zeros((T, N), dtype)
        out["y_next_trace"] = np.zeros((T,) + y.shape, dtype)  # post-transition, PRE-reset
    for s in range(T):
        if trace:
            out["y_trace"][s] = y
            out["t_trace"][s] = t
        obs = E.observation(env_spec, y).T.copy()  # [N, d]
        a_env, a_store, value, logp = act(env_spec, pol_spec, params, obs, noise_action[s])
        y1, t1 = E.transition(env_spec, y, t, a_env)
        sc1 = step_count + 1
        r = E.reward(env_spec, y, a_env, y1)
        term = E.terminal(env_spec, y1)
        trunc = E.truncate(env_spec, sc1)
        done = term | trunc
        if trunc.any():
            v_next = P.value_only(pol_spec, params, E.observation(env_spec, y1).T.copy())
            r = np.where(trunc, r + dt(gamma) * v_next, r)
        if trace:
            out["y_next_trace"][s] = y1
        y0 = E.initial_from_uniform(env_spec, reset_uniform[s].T, dtype)
        y = np.where(done[None, :], y0, y1)
        t = np.where(done, dt(0.0), t1)
        step_count = np.where(done, 0, sc1).astype(np.int32)
        out["obs"][s], out["act"][s], out["rew"][s] = obs, a_store, r
        out["done"][s], out["logp"][s], out["val"][s] = done, logp, value
    out["last_value"] = P.value_only(pol_spec, params, E.observation(env_spec, y).T.copy())
    out["y"], out["t"], out["step_count"] = y, t, step_count
    return out
