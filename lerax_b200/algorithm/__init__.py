"""PPO with the reference's constructor, `learn` signature and loop structure.

Reference: algorithm/ppo.py:105-561, algorithm/base_algorithm.py:113-276.  The three loop
nests of the reference (T-step scan under an N-way vmap, reverse GAE scan, epochs x
minibatches of loss/grad/Adam) are three CUDA entry points behind the C ABI:
`lrx_rollout`, `lrx_gae`, `lrx_ppo_update` (or `lrx_ppo_adv_sums`/`lrx_ppo_grad`/
`lrx_ppo_apply` with an NCCL all-reduce in between when `torch.distributed` is initialised
with more than one rank).  Host code only sequences launches; nothing is computed on the CPU.
"""

from __future__ import annotations

import copy
import os

from dataclasses import dataclass
from typing import Any, Callable, Sequence

from .. import _lib as L
from ..buffer import RolloutBuffer
from ..callback import (AbstractCallback, CallbackList, IterationContext, ResetContext, TrainingContext)
from ..env.classic_control import AbstractClassicControlEnv, ClassicControlState
from ..policy import MLPActorCriticPolicy
from ..random import Key, split
from ..space import Box

_MASK64 = (1 << 64) - 1


@dataclass
class AdamState:
    """optax.chain(clip_by_global_norm, inject_hyperparams(adam)) state: mu, nu, count."""

    mu: Any
    nu: Any
    count: Any  # int32 CUDA scalar

    @classmethod
    def init(cls, params):
        import torch

        return cls(torch.zeros_like(params), torch.zeros_like(params),
                   torch.zeros((), dtype=torch.int32, device=params.device))

    def clone(self):
        return AdamState(self.mu.clone(), self.nu.clone(), self.count.clone())


@dataclass
class PPOStepState:
    env_state: ClassicControlState
    policy_state: Any = None
    callback_state: Any = None


@dataclass
class PPOState:
    iteration_count: int
    step_state: PPOStepState
    env: AbstractClassicControlEnv
    policy: MLPActorCriticPolicy
    opt_state: AdamState
    callback_state: Any
    # device-side RNG bookkeeping (Philox counters), not present in the reference
    rollout_seed: int = 0
    steps_done: int = 0

    def with_callback_states(self, callback_state):
        self.callback_state = callback_state
        return self


class AbstractAlgorithm:
    pass


def _dist():
    """(rank, world_size, torch.distributed or None)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_rank(), dist.get_world_size(), dist
    return 0, 1, None


class PPO(AbstractAlgorithm):
    """Proximal Policy Optimization; kwargs and defaults of algorithm/ppo.py:160-176."""

    def __init__(self, *, num_envs: int = 4, num_steps: int = 2048, num_epochs: int = 10, num_batches: int = 32,
                 gae_lambda: float = 0.95, gamma: float = 0.99, clip_coefficient: float = 0.2,
                 clip_value_loss: bool = False, entropy_loss_coefficient: float = 0.0,
                 value_loss_coefficient: float = 0.5, max_grad_norm: float = 0.5,
                 normalize_advantages: bool = True, learning_rate: float | Callable[[int], float] = 3e-4):
        self.gamma = gamma
        self.gae_lambda = gae_lambda
        self.num_envs = num_envs
        self.num_steps = num_steps
        self.num_epochs = num_epochs
        self.num_batches = num_batches
        self.batch_size = (self.num_steps * self.num_envs) // num_batches  # ppo.py:183
        self.clip_coefficient = clip_coefficient
        self.clip_value_loss = clip_value_loss
        self.entropy_loss_coefficient = entropy_loss_coefficient
        self.value_loss_coefficient = value_loss_coefficient
        self.max_grad_norm = max_grad_norm
        self.normalize_advantages = normalize_advantages
        self.learning_rate = learning_rate
        self.adam_b1, self.adam_b2, self.adam_eps = 0.9, 0.999, 1e-8  # optax.adam defaults
        self._surrogate = 0   # LrxPpoHyper.surrogate: PPO's clipped ratio
        self._shuffle = True  # PPO permutes the buffer every epoch (base_buffer.py:85-88)
        self._log_keys = ("approx_kl", "loss", "policy_loss", "value_loss", "entropy_loss", "explained_variance")
        self._bufs = None
        if self.batch_size <= 0:
            raise ValueError("num_batches larger than num_envs * num_steps")
        # base_buffer.py:85-88 trims the permutation to a multiple of batch_size and reshapes to
        # (-1, batch_size): the number of minibatches is total // batch_size (>= num_batches)
        self.minibatches_per_epoch = (self.num_steps * self.num_envs) // self.batch_size

    # ------------------------------------------------------------------ helpers
    def num_iterations(self, total_timesteps: int) -> int:
        return total_timesteps // (self.num_envs * self.num_steps)  # ppo.py:196-197

    def consolidate_callbacks(self, callback) -> AbstractCallback:
        # base_algorithm.py:193-206
        if callback is None:
            callback = CallbackList(callbacks=[])
        elif isinstance(callback, Sequence):
            callback = CallbackList(callbacks=list(callback))
        elif not isinstance(callback, AbstractCallback):
            raise TypeError(f"Invalid callback type: {type(callback)}")
        if callback._has_custom_on_step():
            raise NotImplementedError(
                "callbacks with a custom on_step cannot run inside the fused CUDA rollout; "
                "derive per-step statistics from the (reward, done) buffer in on_iteration instead"
            )
        return callback

    def _hyper(self, lr: float) -> L.LrxPpoHyper:
        return L.LrxPpoHyper(self.clip_coefficient, self.value_loss_coefficient, self.entropy_loss_coefficient,
                             self.max_grad_norm, lr, self.adam_b1, self.adam_b2, self.adam_eps,
                             int(self.clip_value_loss), int(self.normalize_advantages), int(self._surrogate))

    def _lr_at(self, count: int) -> float:
        return float(self.learning_rate(count)) if callable(self.learning_rate) else float(self.learning_rate)

    def _exchange(self, policy, dist):
        """The peer-memory gradient exchange of this process (None: single GPU, or NCCL fallback)."""
        if dist is None or getattr(self, "_force_nccl", False):
            return None
        n = policy.n_params + L.LRX_PPO_NSTATS
        if getattr(self, "_ex_key", None) != (n, dist.get_world_size()):
            from .._p2p import make_exchange

            self._ex = make_exchange(n, dist, policy.params.device)
            self._ex_key = (n, dist.get_world_size())
        return self._ex

    def _local_envs(self) -> int:
        rank, world, _ = _dist()
        if self.num_envs % world:
            raise ValueError(f"num_envs={self.num_envs} must be divisible by the world size {world}")
        return self.num_envs // world

    def _buffers(self, env, policy, device):
        """Persistent device buffers for one rollout of the local env shard."""
        torch = L.require_cuda()
        N, T, d = self._local_envs(), self.num_steps, policy.obs_dim
        key = (N, T, d, policy.n_params, str(device), isinstance(env.action_space, Box))
        if self._bufs is not None and self._bufs["key"] == key:
            return self._bufs
        f32 = dict(dtype=torch.float32, device=device)
        rank, world, _ = _dist()
        B_local = self.batch_size // world
        ws_bytes = int(L.lib.lrx_ppo_workspace_bytes(policy.desc(), max(B_local, 1)))
        self._bufs = dict(
            key=key,
            obs=torch.empty((T, N, d), **f32),
            act=torch.empty((T, N), dtype=torch.float32 if isinstance(env.action_space, Box) else torch.int32,
                            device=device),
            rew=torch.empty((T, N), **f32), done=torch.empty((T, N), dtype=torch.uint8, device=device),
            logp=torch.empty((T, N), **f32), val=torch.empty((T, N), **f32),
            last_value=torch.empty((N,), **f32),
            ev_sums=torch.zeros((L.LRX_GAE_EV_DOUBLES,), dtype=torch.float64, device=device),
            stats=torch.zeros((self.num_epochs, self.minibatches_per_epoch, L.LRX_PPO_NSTATS), **f32),
            nonfinite=torch.zeros((), dtype=torch.int32, device=device),
            workspace=torch.empty((ws_bytes,), dtype=torch.uint8, device=device),
            gradbuf=torch.zeros((policy.n_params + L.LRX_PPO_NSTATS,), **f32),
            adv_sums=torch.zeros((self.minibatches_per_epoch, 2), dtype=torch.float64, device=device),
            # row-major copy of the buffer for the minibatch gather of the tensor-core update kernel (default widths)
            packed=(torch.empty((T * N, L.LRX_PACKED_FLOATS), **f32) if (d <= 8 and policy.is_default_widths()) else None),
            # activations of the step-by-step (GEMM-batched) rollout of non-default widths
            rollout_ws=(None if policy.is_default_widths() else torch.empty(
                (max(int(L.lib.lrx_rollout_workspace_bytes(policy.desc(), N)), 1),), dtype=torch.uint8, device=device)),
        )
        return self._bufs

    # ------------------------------------------------------------------ reset
    def reset(self, env, policy, *, key: Key | int, callback: AbstractCallback) -> PPOState:
        """ppo.py:318-344: initial env states for all envs, fresh optimiser state."""
        torch = L.require_cuda()
        step_key, callback_key, rollout_key = split(key, 3)
        if not policy.params.is_cuda:
            policy = policy.to(torch.device("cuda", torch.cuda.current_device()))  # H2D of the parameters
        policy = policy.with_params(policy.params.clone())  # learn() does not mutate its input
        env = copy.copy(env)  # curriculum callbacks rewrite fields of the TRAINING state's env, not the caller's
        rank, world, _ = _dist()
        N = self._local_envs()
        dev = policy.params.device
        st = ClassicControlState(
            torch.empty((env.state_dim, N), dtype=torch.float32, device=dev),
            torch.empty((N,), dtype=torch.float32, device=dev),
            torch.empty((N,), dtype=torch.int32, device=dev),
        )
        rng = L.LrxRng(int(step_key) & _MASK64, rank * N, 0)
        L.check(L.lib.lrx_env_reset(env.desc(), st.c_state(), N, rng, None, L.stream_ptr()), "lrx_env_reset")
        ctx = ResetContext(locals())
        step_cb = callback.step_reset(ctx, key=callback_key)
        cb_state = callback.reset(ctx, key=callback_key)
        return PPOState(0, PPOStepState(st, None, step_cb), env, policy, AdamState.init(policy.params), cb_state,
                        rollout_seed=int(rollout_key) & _MASK64, steps_done=0)

    # ------------------------------------------------------------------ rollout
    def collect_rollout(self, env, policy, step_state: PPOStepState, callback, key=None, *, seed: int = 0,
                        step_offset: int = 0, replay=None) -> tuple[PPOStepState, RolloutBuffer]:
        """ppo.py:290-316 for all local envs at once: fused rollout kernel, then GAE."""
        rank, world, _ = _dist()
        N, T = self._local_envs(), self.num_steps
        b = self._buffers(env, policy, policy.params.device)
        out = L.LrxRolloutOut(L.ptr(b["obs"]), L.ptr(b["act"]), L.ptr(b["rew"]), L.ptr(b["done"]), L.ptr(b["logp"]),
                              L.ptr(b["val"]), L.ptr(b["last_value"]), None, None, None, L.ptr(b["rollout_ws"]),
                              0 if b["rollout_ws"] is None else b["rollout_ws"].numel())
        rng = L.LrxRng(seed & _MASK64, rank * N, step_offset & 0xFFFFFFFF)
        L.check(L.lib.lrx_rollout(env.desc(), policy.desc(), L.ptr(policy.params), step_state.env_state.c_state(),
                                  rng, replay, out, N, T, float(self.gamma), L.stream_ptr()), "lrx_rollout")
        # the T on_step calls per env of ppo.py:262-277, for the callbacks whose step state the fused path supports
        step_state.callback_state = callback.advance_step_state(step_state.callback_state, b["rew"], b["done"], N, T)
        buf = RolloutBuffer(b["obs"], b["act"], b["rew"], b["done"], b["logp"], b["val"],
                            squeeze_env=(self.num_envs == 1))
        buf = buf.compute_returns_and_advantages(b["last_value"], self.gae_lambda, self.gamma, ev_sums=b["ev_sums"])
        return step_state, buf

    # ------------------------------------------------------------------ train
    def train(self, policy, opt_state: AdamState, buffer: RolloutBuffer, *, key: Key | int, applied: int = 0):
        """ppo.py:530-561: num_epochs x num_batches minibatch updates; returns the log dict."""
        torch = L.require_cuda()
        rank, world, dist = _dist()
        b = self._bufs
        assert b is not None, "collect_rollout must run before train"
        pol_d = policy.desc()
        view = buffer.view(packed=b.get("packed"))
        B_global = self.batch_size
        B_local = B_global // world
        if B_local * world != B_global:
            raise ValueError(f"batch_size={B_global} must be divisible by the world size {world}")
        stats = b["stats"]
        ws, ws_bytes = L.ptr(b["workspace"]), b["workspace"].numel()
        epoch_keys = split(key, self.num_epochs)
        count = applied
        for e in range(self.num_epochs):
            # base_buffer.py:85-88.  With several ranks each rank permutes its own shard and
            # contributes B/world rows to every minibatch (SURVEY.md 8e mode ii).
            perm_key = Key((int(epoch_keys[e]) + rank) & _MASK64) if self._shuffle else None
            idx = buffer.batch_indices(B_local, key=perm_key)[: self.minibatches_per_epoch]
            nb = idx.shape[0]
            if world == 1:
                # one C call per epoch; a callable learning_rate (optax schedule of the update count, ppo.py:176,192-194)
                # travels as one rate per minibatch
                h = self._hyper(self._lr_at(count))
                lrs = None
                if callable(self.learning_rate):
                    import ctypes

                    lrs = (ctypes.c_float * nb)(*[self._lr_at(count + i) for i in range(nb)])
                L.check(L.lib.lrx_ppo_update_schedule(pol_d, L.ptr(policy.params), L.ptr(opt_state.mu), L.ptr(opt_state.nu),
                                                      L.ptr(opt_state.count), view, L.ptr(idx), nb, B_local, h, lrs,
                                                      L.ptr(stats[e]), L.ptr(b["nonfinite"]), ws, ws_bytes, L.stream_ptr()),
                        "lrx_ppo_update")
                count += nb
                continue
            L.check(L.lib.lrx_ppo_adv_sums(view.adv, view.N, view.T, L.ptr(idx), nb, B_local, L.ptr(b["adv_sums"]),
                                           L.stream_ptr()), "lrx_ppo_adv_sums")
            if dist is not None:
                dist.all_reduce(b["adv_sums"])  # one collective for the whole epoch
            ex = self._exchange(policy, dist)
            if ex is not None and idx.is_contiguous():
                # the whole epoch in one C call (lrx_ppo_update_p2p): per minibatch the gradient kernel and ONE cooperative
                # kernel for reduce + NVLink exchange + clip + Adam
                lrs = None
                if callable(self.learning_rate):
                    import ctypes

                    lrs = (ctypes.c_float * nb)(*[self._lr_at(count + i) for i in range(nb)])
                ex.update_epoch(pol_d, policy, opt_state, view, idx, B_local, B_global, b["adv_sums"],
                                self._hyper(self._lr_at(count)), lrs, stats[e], b["nonfinite"], b["workspace"])
                count += nb
                continue
            for i in range(nb):
                h = self._hyper(self._lr_at(count))
                if ex is not None:
                    # the local gradient goes straight into this rank's exchange slot, published as soon as it is complete
                    ex.grad(pol_d, policy, view, idx[i], B_local, B_global, b["adv_sums"][i], h, b["workspace"])
                else:
                    L.check(L.lib.lrx_ppo_grad(pol_d, L.ptr(policy.params), view, L.ptr(idx[i]), B_local, B_global,
                                               L.ptr(b["adv_sums"][i]), h, L.ptr(b["gradbuf"]), ws, ws_bytes,
                                               L.stream_ptr()), "lrx_ppo_grad")
                if ex is not None:
                    # SUM over the ranks' exchange slots (NVLink peer memory, rank order) + clip + Adam: one launch
                    ex.apply(pol_d, policy, opt_state, h, stats[e, i], b["nonfinite"])
                    count += 1
                    continue
                if dist is not None:
                    dist.all_reduce(b["gradbuf"])  # grads || stat sums, SUM over ranks (NCCL)
                L.check(L.lib.lrx_ppo_apply(pol_d, L.ptr(policy.params), L.ptr(opt_state.mu), L.ptr(opt_state.nu),
                                            L.ptr(opt_state.count), L.ptr(b["gradbuf"]), h, L.ptr(stats[e, i]),
                                            L.ptr(b["nonfinite"]), L.stream_ptr()), "lrx_ppo_apply")
                count += 1

        # stats: mean over minibatches, then over epochs (ppo.py:520,551)
        s = stats.mean(dim=1).mean(dim=0)
        ev = b["ev_sums"][:4].clone()
        if dist is not None:
            dist.all_reduce(ev)
        n_tot = float(self.num_envs * self.num_steps)
        var_ret = ev[1] / n_tot - (ev[0] / n_tot) ** 2
        var_err = ev[3] / n_tot - (ev[2] / n_tot) ** 2
        eps = 1.1920928955078125e-07
        log = {
            "approx_kl": s[0], "loss": s[1], "policy_loss": s[2], "value_loss": s[3], "entropy_loss": s[4],
            "explained_variance": (1.0 - var_err / (var_ret + eps)).to(torch.float32),  # ppo.py:523-528
        }
        return policy, opt_state, {k: v for k, v in log.items() if k in self._log_keys}

    # ------------------------------------------------------------------ iteration
    def iteration(self, state: PPOState, *, key: Key | int, callback: AbstractCallback) -> PPOState:
        """ppo.py:346-394."""
        rollout_key, train_key, callback_key = split(key, 3)
        step_state, buffer = self.collect_rollout(state.env, state.policy, state.step_state, callback, rollout_key,
                                                  seed=state.rollout_seed, step_offset=state.steps_done)
        applied = state.iteration_count * self.num_epochs * self.minibatches_per_epoch
        policy, opt_state, log = self.train(state.policy, state.opt_state, buffer, key=train_key, applied=applied)
        state.step_state, state.policy, state.opt_state = step_state, policy, opt_state
        state.iteration_count += 1
        state.steps_done += self.num_steps
        state.last_buffer = buffer
        cb_state = callback.on_iteration(
            IterationContext(state.callback_state, state.step_state.callback_state, state.env, state.policy,
                             state.iteration_count, state.opt_state, log, self, locals()),
            key=callback_key,
        )
        state = state.with_callback_states(cb_state)
        state, new_cb = callback.apply_curriculum(state, state.callback_state)
        return state.with_callback_states(new_cb)

    def check_finite(self):
        """The eqx.error_if of ppo.py:413-417, surfaced once per call (one host sync)."""
        if self._bufs is not None and int(self._bufs["nonfinite"].item()) != 0:
            self._bufs["nonfinite"].zero_()
            raise FloatingPointError("Non-finite values / log_probs / entropy in ppo_loss (update skipped).")
        if getattr(self, "_ex", None) is not None:
            self._ex.check()

    # ------------------------------------------------------------------ learn
    def learn(self, env, policy, total_timesteps: int, *, key: Key | int, callback=None):
        """base_algorithm.py:208-276.  Returns the trained policy (parameters on the GPU)."""
        L.require_cuda()
        callback_start_key, reset_key, learn_key, callback_end_key = split(key, 4)
        callback = self.consolidate_callbacks(callback)
        state = self.reset(env, policy, key=reset_key, callback=callback)
        state = state.with_callback_states(callback.on_training_start(
            TrainingContext(state.callback_state, state.step_state.callback_state, env, state.policy,
                            total_timesteps, state.iteration_count, state.opt_state, self, locals()),
            key=callback_start_key))
        n_iter = self.num_iterations(total_timesteps)
        has_callbacks = not (isinstance(callback, CallbackList) and not callback.callbacks)
        for k in split(learn_key, n_iter) if n_iter else []:
            state = self.iteration(state, key=k, callback=callback)
            if has_callbacks:
                self.check_finite()
        self.check_finite()
        state = state.with_callback_states(callback.on_training_end(
            TrainingContext(state.callback_state, state.step_state.callback_state, env, state.policy,
                            total_timesteps, state.iteration_count, state.opt_state, self, locals()),
            key=callback_end_key))
        self.last_state = state
        return state.policy


class A2C(PPO):
    """Advantage Actor-Critic; kwargs and defaults of algorithm/a2c.py:150-177.

    The rollout step and `collect_rollout` of the reference's A2C are line-for-line those of PPO
    (a2c.py:183-300 vs ppo.py:199-316), so the same rollout and GAE kernels serve it; `train`
    (a2c.py:415-444) is ONE gradient step on the whole buffer with the policy loss
    -mean(log_prob * advantage) (a2c.py:381-411): the update kernels with `surrogate = 1`,
    one minibatch of N*T rows in buffer order."""

    def __init__(self, *, num_envs: int = 4, num_steps: int = 5, gae_lambda: float = 1.0, gamma: float = 0.99,
                 entropy_loss_coefficient: float = 0.0, value_loss_coefficient: float = 0.5,
                 max_grad_norm: float = 0.5, normalize_advantages: bool = False,
                 learning_rate: float | Callable[[int], float] = 7e-4):
        super().__init__(num_envs=num_envs, num_steps=num_steps, num_epochs=1, num_batches=1, gae_lambda=gae_lambda,
                         gamma=gamma, entropy_loss_coefficient=entropy_loss_coefficient,
                         value_loss_coefficient=value_loss_coefficient, max_grad_norm=max_grad_norm,
                         normalize_advantages=normalize_advantages, learning_rate=learning_rate)
        self._surrogate = 1
        self._shuffle = False
        self._log_keys = ("loss", "policy_loss", "value_loss", "entropy_loss")  # a2c.py:438-443


class REINFORCE(A2C):
    """Monte-Carlo policy gradient with a value baseline; kwargs and defaults of
    algorithm/reinforce.py:145-168 (GAE with lambda fixed to 1, no entropy term, :372-396)."""

    def __init__(self, *, num_envs: int = 1, num_steps: int = 512, gamma: float = 0.99,
                 normalize_advantages: bool = True, value_loss_coefficient: float = 0.5, max_grad_norm: float = 0.5,
                 learning_rate: float | Callable[[int], float] = 3e-4):
        super().__init__(num_envs=num_envs, num_steps=num_steps, gae_lambda=1.0, gamma=gamma,
                         entropy_loss_coefficient=0.0, value_loss_coefficient=value_loss_coefficient,
                         max_grad_norm=max_grad_norm, normalize_advantages=normalize_advantages,
                         learning_rate=learning_rate)
        self._log_keys = ("loss", "policy_loss", "value_loss")  # reinforce.py:425-429


__all__ = ["A2C", "AbstractAlgorithm", "AdamState", "PPO", "PPOState", "PPOStepState", "REINFORCE"]
