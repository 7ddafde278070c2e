"""ProgressBarCallback (callback/progress_bar.py:120-200): env steps done so far.

The reference counts steps in `on_step` (:177-181); the fused rollout advances the per-env counters
by T once per rollout.  The bar itself is tqdm (rich's live display is host-side I/O)."""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any

from .. import _lib as L
from . import AbstractCallback, AbstractCallbackStepState, EmptyCallbackState


@dataclass
class ProgressBarCallbackStepState(AbstractCallbackStepState):
    steps: Any  # int64 [N]


class ProgressBarCallback(AbstractCallback):
    _fused_on_step = True

    def __init__(self, total_timesteps: int, name: str | None = None, env=None, policy=None, *, disable: bool = False):
        if name is None:
            who = getattr(policy, "name", None)
            where = getattr(env, "name", None)
            name = "Training" + (f" {who}" if who else "") + (f" on {where}" if where else "")
        self.name = name
        self.total = int(total_timesteps)
        self.completed = 0
        self._bar = None
        self._disable = disable

    def reset(self, ctx, *, key=None):
        if not self._disable:
            from tqdm import tqdm

            self._bar = tqdm(total=self.total, desc=self.name, unit="step")
        return EmptyCallbackState()

    def step_reset(self, ctx, *, key=None):
        torch = L.require_cuda()
        return ProgressBarCallbackStepState(torch.zeros((int(ctx.locals["N"]),), dtype=torch.int64,
                                                        device=ctx.locals["dev"]))

    def on_step(self, ctx, *, key=None):
        return ProgressBarCallbackStepState(ctx.state.steps + 1)

    def advance_step_state(self, step_state, rewards, dones, num_envs, num_steps):
        step_state.steps += int(num_steps)
        return step_state

    def on_iteration(self, ctx, *, key=None):
        completed = int(ctx.step_state.steps.sum().item())
        if self._bar is not None:
            self._bar.update(completed - self.completed)
        self.completed = completed
        return ctx.state

    def stop(self) -> None:
        if self._bar is not None:
            self._bar.close()
            self._bar = None
