"""TimeLimit wrapper (reference: wrapper/misc.py:115-206).

The step counter lives in the env state (`step_count`) and the truncation test
`step_count >= max_episode_steps` runs inside the rollout kernel, where it triggers the
value bootstrap of algorithm/ppo.py:248-259.
"""

from __future__ import annotations

import copy


class TimeLimit:
    def __new__(cls, env, max_episode_steps: int):
        if int(max_episode_steps) <= 0:
            raise ValueError("max_episode_steps must be positive")
        wrapped = copy.copy(env)
        wrapped.max_episode_steps = int(max_episode_steps)
        wrapped.env = env
        return wrapped


__all__ = ["TimeLimit"]
