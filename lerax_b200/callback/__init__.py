"""Callback protocol of the training loop (reference: callback/base_callback.py:30-185,
callback/list.py:48-162, callback/empty.py:17).

The hooks run on the host between kernel launches.  `on_step` is special: in the reference
it is traced into the per-step scan; here the rollout is one fused kernel, so only callbacks
whose `on_step` state is a pure function of the (reward, done) streams can be supported:
they implement `advance_step_state`, called once per rollout with the [T][N] reward / done
buffers (LoggingCallback -> the `lrx_episode_stats` kernel, ProgressBarCallback -> a counter);
a user callback that overrides `on_step` with anything else raises NotImplementedError.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any


class AbstractCallbackStepState:
    pass


class AbstractCallbackState:
    pass


@dataclass
class ResetContext:
    locals: dict = field(default_factory=dict)


@dataclass
class StepContext:
    state: Any
    env: Any
    policy: Any
    done: Any
    reward: Any
    locals: dict = field(default_factory=dict)


@dataclass
class IterationContext:
    state: Any
    step_state: Any
    env: Any
    policy: Any
    iteration_count: Any
    opt_state: Any
    training_log: dict
    algorithm: Any
    locals: dict = field(default_factory=dict)


@dataclass
class TrainingContext:
    state: Any
    step_state: Any
    env: Any
    policy: Any
    total_timesteps: int
    iteration_count: Any
    opt_state: Any
    algorithm: Any
    locals: dict = field(default_factory=dict)


class AbstractCallback:
    """Seven hooks + apply_curriculum, as in callback/base_callback.py:119-185."""

    def reset(self, ctx: ResetContext, *, key=None):
        return None

    def step_reset(self, ctx: ResetContext, *, key=None):
        return None

    def on_step(self, ctx: StepContext, *, key=None):
        return ctx.state

    def on_iteration(self, ctx: IterationContext, *, key=None):
        return ctx.state

    def on_training_start(self, ctx: TrainingContext, *, key=None):
        return ctx.state

    def on_training_end(self, ctx: TrainingContext, *, key=None):
        return ctx.state

    def continue_training(self, ctx: IterationContext, *, key=None):
        return True

    def apply_curriculum(self, state, callback_state):
        return state, callback_state

    # The fused-rollout replacement of T on_step calls per env: rewards / dones are the [T][N] device buffers
    # of the rollout that just ran.  Built-in callbacks whose step state is a function of those streams override
    # this (and set _fused_on_step); the default keeps the state.
    _fused_on_step = False

    def advance_step_state(self, step_state, rewards, dones, num_envs: int, num_steps: int):
        return step_state

    # whether on_step is the inherited no-op (anything else cannot run inside the fused rollout)
    def _has_custom_on_step(self) -> bool:
        return type(self).on_step is not AbstractCallback.on_step and not self._fused_on_step


class EmptyCallbackState(AbstractCallbackState):
    """callback/base_callback.py: the state of callbacks that keep none."""


class EmptyCallbackStepState(AbstractCallbackStepState):
    pass


class EmptyCallback(AbstractCallback):
    """callback/empty.py:17: every hook hands its (empty) state back unchanged."""

    def reset(self, ctx, *, key=None):
        return EmptyCallbackState()

    def step_reset(self, ctx, *, key=None):
        return EmptyCallbackStepState()


class CallbackList(AbstractCallback):
    def __init__(self, callbacks):
        self.callbacks = list(callbacks)

    def reset(self, ctx, *, key=None):
        return [cb.reset(ctx, key=key) for cb in self.callbacks]

    def step_reset(self, ctx, *, key=None):
        return [cb.step_reset(ctx, key=key) for cb in self.callbacks]

    def _each(self, hook, ctx, key):
        states = ctx.state if isinstance(ctx.state, list) else [None] * len(self.callbacks)
        step_states = getattr(ctx, "step_state", None)
        step_states = step_states if isinstance(step_states, list) else [None] * len(self.callbacks)
        out = []
        for cb, st, sst in zip(self.callbacks, states, step_states):
            fields = {**ctx.__dict__, "state": st}
            if "step_state" in fields:
                fields["step_state"] = sst  # every callback sees its own states (callback/list.py:96-128)
            out.append(getattr(cb, hook)(type(ctx)(**fields), key=key))
        return out

    def on_iteration(self, ctx, *, key=None):
        return self._each("on_iteration", ctx, key)

    def on_training_start(self, ctx, *, key=None):
        return self._each("on_training_start", ctx, key)

    def on_training_end(self, ctx, *, key=None):
        return self._each("on_training_end", ctx, key)

    def continue_training(self, ctx, *, key=None):
        return all(self._each("continue_training", ctx, key)) if self.callbacks else True

    def advance_step_state(self, step_state, rewards, dones, num_envs, num_steps):
        states = step_state if isinstance(step_state, list) else [None] * len(self.callbacks)
        return [cb.advance_step_state(st, rewards, dones, num_envs, num_steps)
                for cb, st in zip(self.callbacks, states)]

    def apply_curriculum(self, state, callback_state):
        states = callback_state if isinstance(callback_state, list) else [None] * len(self.callbacks)
        new = []
        for cb, st in zip(self.callbacks, states):
            state, st = cb.apply_curriculum(state, st)
            new.append(st)
        return state, new

    def _has_custom_on_step(self) -> bool:
        return any(cb._has_custom_on_step() for cb in self.callbacks)


from .logging import (AbstractLoggingBackend, ConsoleBackend, HistoryBackend, LoggingCallback,  # noqa: E402
                      LoggingCallbackStepState)
from .progress_bar import ProgressBarCallback, ProgressBarCallbackStepState  # noqa: E402
from .tensorboard import TensorBoardBackend  # noqa: E402

__all__ = ["AbstractCallback", "AbstractCallbackState", "AbstractCallbackStepState", "AbstractLoggingBackend",
           "CallbackList", "ConsoleBackend", "EmptyCallback", "EmptyCallbackState", "EmptyCallbackStepState",
           "HistoryBackend", "IterationContext", "LoggingCallback", "LoggingCallbackStepState", "ProgressBarCallback",
           "ProgressBarCallbackStepState", "ResetContext", "StepContext", "TensorBoardBackend", "TrainingContext"]
