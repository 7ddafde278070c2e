"""Evaluation episodes (reference: benchmark/__init__.py:12-126).

`average_reward(env, policy, num_episodes, max_steps, deterministic, key=...)` runs the episodes as ONE batch
through the batched env / policy entry points (lrx_env_reset / lrx_env_observe / lrx_policy_forward /
lrx_env_step): an episode that has ended keeps its state and contributes zero reward, as the `lax.cond` of
`rollout_scan` does (:76-104).  `deterministic=True` takes the mode of the action distribution."""

from __future__ import annotations

from . import _lib as L
from .random import split


def rollout_scan(env, policy, *, key, deterministic: bool = False, max_steps: int = 1024, num_episodes: int = 1):
    """Sum of rewards per episode, float32 [num_episodes] (benchmark/__init__.py:59-105, vmapped)."""
    torch = L.require_cuda()
    init_key, *step_keys = split(key, max_steps + 1)
    state = env.initial(key=init_key, num_envs=num_episodes)
    dev = state.t.device
    done = torch.zeros(num_episodes, dtype=torch.bool, device=dev)
    total = torch.zeros(num_episodes, dtype=torch.float32, device=dev)
    for i, k in enumerate(step_keys):
        obs = env.observation(state)
        _, action = policy(None, obs) if deterministic else policy(None, obs, key=k)
        nxt, reward, term, trunc, _ = env.step_full(state, action)
        total += torch.where(done, torch.zeros_like(reward), reward)
        # finished episodes keep their state (the done branch of the reference returns the carry unchanged)
        keep = done
        nxt.y = torch.where(keep.unsqueeze(0), state.y, nxt.y)
        nxt.t = torch.where(keep, state.t, nxt.t)
        nxt.step_count = torch.where(keep, state.step_count, nxt.step_count)
        done = done | term | trunc
        state = nxt
        if (i & 31) == 31 and bool(done.all()):  # the reference scans all max_steps; the sum cannot change any more
            break
    return total


def average_reward(env, policy, num_episodes: int = 4, max_steps: int | None = 1024, deterministic: bool = False, *,
                   key):
    """Mean episode reward over num_episodes episodes (benchmark/__init__.py:108-126); a Python float."""
    if max_steps is None:
        max_steps = 1 << 20  # rollout_while: until the episode ends
    return float(rollout_scan(env, policy, key=key, deterministic=deterministic, max_steps=max_steps,
                              num_episodes=num_episodes).mean())


__all__ = ["average_reward", "rollout_scan"]
