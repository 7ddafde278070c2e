"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the built-in `on_step` callback states.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg only; the product
(lerax_b200) never imports it.

Reference: `LoggingCallbackStepState.next` (src/lerax/callback/logging/callback.py:139-176),
applied once per env step by `LoggingCallback.on_step` (:477-480) inside the rollout scan
(algorithm/ppo.py:262-277); `ProgressBarCallbackStepState` (callback/progress_bar.py:177-181) is
the `step` counter.  The reference's tests hold no known-answer vectors for these states, so this
restatement is pinned only against a scalar transcription of `next` (tests/test_oracle_callback.py):
parity unpinned against the reference itself.
"""
from __future__ import annotations

import numpy as np


def initial(n_envs: int) -> dict:
    """LoggingCallbackStepState.initial() for n_envs envs (callback.py:127-137)."""
    return {
        "step": np.zeros(n_envs, np.int64),
        "episode_return": np.zeros(n_envs, np.float32),
        "episode_length": np.zeros(n_envs, np.int32),
        "episode_done": np.zeros(n_envs, bool),
        "average_return": np.zeros(n_envs, np.float32),
        "average_length": np.zeros(n_envs, np.float32),
    }


def next_state(state: dict, reward, done, alpha: float) -> dict:
    """One env step for all envs at once (callback.py:139-176); float32 like JAX without x64."""
    f = np.float32
    a, one_minus_a = f(alpha), f(1.0) - f(alpha)
    prev_done = state["episode_done"]
    episode_return = state["episode_return"] * (f(1.0) - prev_done.astype(f)) + reward.astype(f)
    episode_length = state["episode_length"] * (1 - prev_done.astype(np.int32)) + 1
    average_return = np.where(done, a * episode_return + one_minus_a * state["average_return"],
                              state["average_return"]).astype(f)
    average_length = np.where(done, a * episode_length.astype(f) + one_minus_a * state["average_length"],
                              state["average_length"]).astype(f)
    return {
        "step": state["step"] + 1,
        "episode_return": episode_return.astype(f),
        "episode_length": episode_length.astype(np.int32),
        "episode_done": np.asarray(done, bool).copy(),
        "average_return": average_return,
        "average_length": average_length,
    }


def scan(state: dict, rewards_tn, dones_tn, alpha: float) -> dict:
    """Advance over a [T][N] rollout (the per-step on_step calls of one collect_rollout)."""
    for t in range(rewards_tn.shape[0]):
        state = next_state(state, rewards_tn[t], dones_tn[t], alpha)
    return state


def iteration_scalars(state: dict) -> dict:
    """What LoggingCallback.on_iteration derives from the step states (callback.py:497-501)."""
    return {
        "episode/return": float(state["average_return"].astype(np.float32).mean(dtype=np.float32)),
        "episode/length": float(state["average_length"].astype(np.float32).mean(dtype=np.float32)),
        "step": int(state["step"].sum()),
    }


def adaptive_curriculum_step_state(metric0, step_values, dones):
    """AbstractAdaptiveCurriculum.on_step scanned over T (curriculum/adaptive.py:64-70), per env:
    episode_metric = where(done, 0, episode_metric + metric_fn(done, reward, locals)).  step_values, dones: [T, N]."""
    m = np.asarray(metric0, np.float32).copy()
    for t in range(step_values.shape[0]):
        m = np.where(dones[t], np.float32(0.0), m + step_values[t].astype(np.float32)).astype(np.float32)
    return m
