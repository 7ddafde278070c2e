"""LoggingCallback: episode-return / episode-length EMAs and training scalars per iteration.

Reference: callback/logging/callback.py (`LoggingCallbackStepState` :90-176, `LoggingCallback`
:380-560) and the backend protocol of callback/logging/backend.py.  In the reference `on_step`
advances the per-env step state inside the rollout scan; here the rollout is one fused kernel and the
same state is advanced once per rollout by `lrx_episode_stats` from the reward / done buffers that
kernel wrote (include/lerax_b200.h).  TensorBoard / W&B backends and video recording are host-side
I/O outside the path; `ConsoleBackend` and `HistoryBackend` implement the protocol.
"""

from __future__ import annotations

from dataclasses import dataclass
from datetime import datetime
from typing import Any, Sequence

from .. import _lib as L
from . import AbstractCallback, AbstractCallbackStepState, EmptyCallbackState


class AbstractLoggingBackend:
    """callback/logging/backend.py: open / log_scalars / log_hparams / on_training_start / close."""

    def open(self, name: str) -> None:
        pass

    def log_scalars(self, scalars: dict, step: int) -> None:
        raise NotImplementedError

    def log_hparams(self, hparams: dict) -> None:
        pass

    def on_training_start(self, total_timesteps: int, total_iterations: int) -> None:
        pass

    def close(self) -> None:
        pass


class HistoryBackend(AbstractLoggingBackend):
    """Keeps every record in memory: `records` is a list of (step, scalars)."""

    def __init__(self):
        self.name = None
        self.records: list[tuple[int, dict]] = []
        self.hparams: dict = {}

    def open(self, name):
        self.name = name

    def log_scalars(self, scalars, step):
        self.records.append((int(step), dict(scalars)))

    def log_hparams(self, hparams):
        self.hparams.update(hparams)


class ConsoleBackend(AbstractLoggingBackend):
    """One line per iteration on stdout (the reference renders a rich table, console.py)."""

    def __init__(self, every: int = 1):
        self.every = max(1, int(every))
        self._n = 0

    def log_scalars(self, scalars, step):
        self._n += 1
        if self._n % self.every:
            return
        body = "  ".join(f"{k}={v:.4g}" for k, v in sorted(scalars.items()))
        print(f"[step {int(step):>12d}] {body}", flush=True)


@dataclass
class LoggingCallbackStepState(AbstractCallbackStepState):
    """Per-env arrays [N] on the GPU; field names as in callback.py:104-109."""

    step: Any            # int64
    episode_return: Any  # float32
    episode_length: Any  # int32
    episode_done: Any    # uint8 (bool)
    average_return: Any  # float32
    average_length: Any  # float32

    @classmethod
    def initial(cls, num_envs: int, device) -> "LoggingCallbackStepState":
        torch = L.require_cuda()
        z = lambda dt: torch.zeros((num_envs,), dtype=dt, device=device)  # noqa: E731
        return cls(z(torch.int64), z(torch.float32), z(torch.int32), z(torch.uint8), z(torch.float32),
                   z(torch.float32))

    def c_struct(self) -> L.LrxEpisodeStats:
        return L.LrxEpisodeStats(L.ptr(self.step), L.ptr(self.episode_return), L.ptr(self.episode_length),
                                 L.ptr(self.episode_done), L.ptr(self.average_return), L.ptr(self.average_length))

    def next(self, reward, done, alpha: float) -> "LoggingCallbackStepState":
        """One env step for every env (callback.py:139-176): the kernel with T = 1, in place."""
        n = self.step.numel()
        rew = reward.to(dtype=self.episode_return.dtype).reshape(1, n).contiguous()
        dn = done.to(dtype=self.episode_done.dtype).reshape(1, n).contiguous()
        L.check(L.lib.lrx_episode_stats(L.ptr(rew), L.ptr(dn), self.c_struct(), n, 1, float(alpha), L.LRX_LAYOUT_TN,
                                        L.stream_ptr()), "lrx_episode_stats")
        return self


def _extract_hparams(obj) -> dict:
    out = {}
    for k, v in vars(obj).items():
        if k.startswith("_"):
            continue
        if isinstance(v, (bool, int, float, str)):
            out[k] = v
    return out


class LoggingCallback(AbstractCallback):
    """Same constructor arguments as callback.py:417-430 (video_* are accepted and must stay disabled)."""

    _fused_on_step = True

    def __init__(self, backend: AbstractLoggingBackend | Sequence[AbstractLoggingBackend], name: str | None = None,
                 env=None, policy=None, alpha: float = 0.9, hparams: dict | None = None, video_interval: int = 0,
                 video_num_steps: int = 128, video_width: int = 640, video_height: int = 480,
                 video_fps: float = 50.0) -> None:
        self._backends = [backend] if isinstance(backend, AbstractLoggingBackend) else list(backend)
        self._hparams = hparams
        self.alpha = float(alpha)
        if video_interval > 0:
            raise NotImplementedError("video recording needs the renderers (render/), which are out of scope")
        if name is None:
            parts = []
            if policy is not None:
                parts.append(getattr(policy, "name", type(policy).__name__))
            if env is not None:
                parts.append(getattr(env, "name", type(env).__name__))
            parts.append(datetime.now().strftime("%Y%m%d_%H%M%S"))
            name = "_".join(parts)
        self._name = name
        for b in self._backends:
            b.open(name)

    def reset(self, ctx, *, key=None):
        return EmptyCallbackState()

    def step_reset(self, ctx, *, key=None):
        loc = ctx.locals
        return LoggingCallbackStepState.initial(int(loc["N"]), loc["dev"])

    def on_step(self, ctx, *, key=None):
        return ctx.state.next(ctx.reward, ctx.done, self.alpha)

    def advance_step_state(self, step_state, rewards, dones, num_envs, num_steps):
        L.check(L.lib.lrx_episode_stats(L.ptr(rewards), L.ptr(dones), step_state.c_struct(), int(num_envs),
                                        int(num_steps), self.alpha, L.LRX_LAYOUT_TN, L.stream_ptr()),
                "lrx_episode_stats")
        return step_state

    def on_iteration(self, ctx, *, key=None):
        torch = L.require_cuda()
        st: LoggingCallbackStepState = ctx.step_state
        # sums over this rank's envs; with several ranks one small all-reduce gives the global means
        sums = torch.stack([st.average_return.sum(dtype=torch.float64), st.average_length.sum(dtype=torch.float64),
                            st.step.sum().to(torch.float64),
                            torch.tensor(float(st.step.numel()), dtype=torch.float64, device=st.step.device)])
        rank = 0
        try:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized():
                dist.all_reduce(sums)
                rank = dist.get_rank()
        except ImportError:  # pragma: no cover
            pass
        ret_sum, len_sum, last_step, n = (float(x) for x in sums.tolist())
        scalars = {f"train/{k}": float(v) for k, v in ctx.training_log.items()}
        lr = getattr(ctx.algorithm, "_lr_at", None)
        applied = int(ctx.opt_state.count.item()) if hasattr(ctx.opt_state, "count") else 0
        scalars["train/learning_rate"] = float(lr(applied)) if lr is not None else float("nan")
        scalars["episode/return"] = ret_sum / n
        scalars["episode/length"] = len_sum / n
        if rank == 0:
            for b in self._backends:
                b.log_scalars(scalars, int(last_step))
        return ctx.state

    def on_training_start(self, ctx, *, key=None):
        total_iterations = ctx.algorithm.num_iterations(ctx.total_timesteps)
        hparams: dict = {}
        hparams.update({f"policy.{k}": v for k, v in _extract_hparams(ctx.policy).items()})
        hparams.update({f"algorithm.{k}": v for k, v in _extract_hparams(ctx.algorithm).items()})
        hparams.update(self._hparams or {})
        for b in self._backends:
            b.on_training_start(ctx.total_timesteps, total_iterations)
            b.log_hparams(hparams)
        return ctx.state

    def close(self) -> None:
        for b in self._backends:
            b.close()
