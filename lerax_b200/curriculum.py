"""Scheduled curricula (reference: curriculum/scheduled.py:23-138).

`ScheduledCurriculum(where=lambda env: env.m, schedule_fn=linear_schedule(...))` rewrites one environment field
after every iteration (`apply_curriculum`, called by `PPO.iteration` after `on_iteration`).  The reference does this
with `eqx.tree_at` on the traced algorithm state; here the env's physical parameters are RUNTIME arguments of every
kernel call (`LrxEnvDesc` is rebuilt from the env object at each launch), so setting the attribute on the training
state's env is all it takes -- nothing is recompiled.  `learn()` works on a copy of the env it was given, so the
caller's object is not modified (the reference is functional).

Adaptive curricula (curriculum/adaptive.py:23-173): `AbstractAdaptiveCurriculum` accumulates a user metric per episode in
its STEP state (`on_step`, :64-70).  The rollout is one fused kernel, so `metric_fn(done, reward, locals)` is evaluated
ONCE per rollout on the whole [T][N] reward / done buffers (it must be elementwise in them -- `reward`, a function of
`reward` and `done` -- and `locals` is empty); the per-env recurrence `m <- 0 if done else m + v` is then replayed over the
T steps in order, so the state equals the reference's step-by-step accumulation bit for bit."""

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


class AdaptiveCurriculumStepState:
    """curriculum/adaptive.py:23-25: the metric accumulated over the current episode, per env ([N] float32)."""

    def __init__(self, episode_metric):
        self.episode_metric = episode_metric


class AdaptiveCurriculumState:
    """curriculum/adaptive.py:28-31."""

    def __init__(self, level, running_metric):
        self.level = level
        self.running_metric = running_metric


class AbstractAdaptiveCurriculum(AbstractCallback):
    """curriculum/adaptive.py:34-118."""

    _fused_on_step = True  # on_step is replayed from the (reward, done) buffers by advance_step_state

    def __init__(self, metric_fn: Callable, smoothing: float = 0.05):
        self.metric_fn = metric_fn
        self.smoothing = smoothing

    def reset(self, ctx, *, key=None):
        return AdaptiveCurriculumState(level=0, running_metric=0.0)

    def step_reset(self, ctx, *, key=None):
        return AdaptiveCurriculumStepState(episode_metric=None)  # zeros [N], allocated at the first rollout

    def advance_step_state(self, step_state, rewards, dones, num_envs: int, num_steps: int):
        """on_step (:64-70) for the T steps of a rollout: v = metric_fn(done, reward, {}); m = where(done, 0, m + v)."""
        import torch

        m = step_state.episode_metric
        if m is None:
            m = torch.zeros((num_envs,), dtype=torch.float32, device=rewards.device)
        done = dones.to(torch.bool)
        v = self.metric_fn(done, rewards, {})
        v = torch.as_tensor(v, dtype=torch.float32, device=rewards.device).expand(rewards.shape)
        zero = torch.zeros((), dtype=torch.float32, device=rewards.device)
        for t in range(num_steps):  # the reference's order of additions, one step at a time
            m = torch.where(done[t], zero, m + v[t])
        return AdaptiveCurriculumStepState(episode_metric=m)

    def on_iteration(self, ctx, *, key=None):
        # :72-82 -- with several envs the running metric is per env, as in the reference (vmapped step states)
        step_metric = ctx.step_state.episode_metric
        running = ctx.state.running_metric
        updated = (1 - self.smoothing) * running + self.smoothing * step_metric
        return AdaptiveCurriculumState(level=ctx.state.level, running_metric=updated)

    def apply_curriculum(self, state, callback_state):
        raise NotImplementedError


class LevelCurriculum(AbstractAdaptiveCurriculum):
    """Advance an env field through discrete levels when the running metric exceeds `threshold`
    (curriculum/adaptive.py:121-173).  With several envs the reference's level / parameter become per-env arrays (its
    step states are vmapped); a kernel launch takes ONE value per physical parameter, so the level advances when the
    MEAN running metric over the envs exceeds the threshold (identical for num_envs = 1)."""

    def __init__(self, where: Callable, levels, metric_fn: Callable, threshold: float, smoothing: float = 0.05):
        super().__init__(metric_fn, smoothing)
        self.where = where
        self.levels = [float(_f32(x)) for x in np.asarray(levels, dtype=np.float64).reshape(-1)]
        self.threshold = threshold
        path = where(_Path())
        if not isinstance(path, _Path) or not path._names:
            raise TypeError("LevelCurriculum: `where` must select an attribute, e.g. lambda env: env.max_speed")
        self._names = path._names

    def apply_curriculum(self, state, callback_state):
        metric = callback_state.running_metric
        metric = float(metric.mean()) if hasattr(metric, "mean") else float(metric)
        level = int(callback_state.level)
        new_level = min(level + 1 if metric > self.threshold else level, len(self.levels) - 1)
        obj = state.env
        for name in self._names[:-1]:
            obj = getattr(obj, name)
        target = obj
        while not hasattr(type(target), self._names[-1]) and self._names[-1] not in vars(target) and hasattr(target, "env"):
            target = target.env
        setattr(target, self._names[-1], self.levels[new_level])
        return state, AdaptiveCurriculumState(level=new_level, running_metric=callback_state.running_metric)


__all__ = ["AbstractAdaptiveCurriculum", "AdaptiveCurriculumState", "AdaptiveCurriculumStepState", "LevelCurriculum",
           "ScheduledCurriculum", "cosine_schedule", "linear_schedule", "step_schedule"]
