"""Scheduled curricula (reference: curriculum/scheduled.py:23-138).

`ScheduledCurriculum(where=lambda env: env.m, schedule_fn=linear_schedule(...))` rewrites one environment field
after every iteration (`apply_curriculum`, called by `PPO.iteration` after `on_iteration`).  The reference does this
with `eqx.tree_at` on the traced algorithm state; here the env's physical parameters are RUNTIME arguments of every
kernel call (`LrxEnvDesc` is rebuilt from the env object at each launch), so setting the attribute on the training
state's env is all it takes -- nothing is recompiled.  `learn()` works on a copy of the env it was given, so the
caller's object is not modified (the reference is functional).  The adaptive curricula's user-defined per-step metric
(curriculum/adaptive.py) cannot run inside the fused rollout and is not provided."""

from __future__ import annotations

import math
from typing import Callable

import numpy as np

from .callback import AbstractCallback, EmptyCallbackState

_f32 = np.float32


def linear_schedule(start: float, end: float, total: int) -> Callable[[int], float]:
    """start -> end over `total` iterations, clamped outside (scheduled.py:23-45)."""
    s, e = _f32(start), _f32(end)

    def schedule(iteration) -> float:
        t = _f32(np.clip(_f32(int(iteration)) / _f32(total), 0.0, 1.0))
        return float(s + (e - s) * t)

    return schedule


def step_schedule(values: list[float], boundaries: list[int]) -> Callable[[int], float]:
    """values[i] from boundaries[i-1] (inclusive) on (scheduled.py:48-70; searchsorted side="right")."""
    if len(values) != len(boundaries) + 1:
        raise ValueError("step_schedule: len(values) must be len(boundaries) + 1")
    vals = [float(_f32(v)) for v in values]
    bounds = [int(b) for b in boundaries]

    def schedule(iteration) -> float:
        return vals[int(np.searchsorted(bounds, int(iteration), side="right"))]

    return schedule


def cosine_schedule(start: float, end: float, total: int) -> Callable[[int], float]:
    """Cosine annealing start -> end over `total` iterations (scheduled.py:73-95)."""
    s, e = _f32(start), _f32(end)

    def schedule(iteration) -> float:
        t = _f32(np.clip(_f32(int(iteration)) / _f32(total), 0.0, 1.0))
        return float(e + (s - e) * (_f32(1.0) + _f32(math.cos(math.pi * float(t)))) / _f32(2.0))

    return schedule


class _Path:
    """Records the attribute path a `where` selector walks (the role of eqx.tree_at's selector)."""

    def __init__(self, names=()):
        object.__setattr__(self, "_names", tuple(names))

    def __getattr__(self, name):
        return _Path(self._names + (name,))


class ScheduledCurriculum(AbstractCallback):
    """Modify an environment field on a fixed schedule (scheduled.py:98-138)."""

    def __init__(self, where: Callable, schedule_fn: Callable[[int], float]):
        self.where = where
        self.schedule_fn = schedule_fn
        path = where(_Path())
        if not isinstance(path, _Path) or not path._names:
            raise TypeError("ScheduledCurriculum: `where` must select an attribute, e.g. lambda env: env.m")
        self._names = path._names

    def reset(self, ctx, *, key=None):
        return EmptyCallbackState()

    def apply_curriculum(self, state, callback_state):
        value = self.schedule_fn(state.iteration_count)
        obj = state.env
        for name in self._names[:-1]:
            obj = getattr(obj, name)
        target = obj
        # a wrapper (TimeLimit) forwards reads to the env it wraps: write where the field lives
        while not hasattr(type(target), self._names[-1]) and self._names[-1] not in vars(target) and hasattr(target, "env"):
            target = target.env
        setattr(target, self._names[-1], value)
        return state, callback_state


__all__ = ["ScheduledCurriculum", "cosine_schedule", "linear_schedule", "step_schedule"]
