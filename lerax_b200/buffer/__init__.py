"""RolloutBuffer (reference: buffer/rollout.py:17-102, buffer/base_buffer.py:38-103).

Storage is time-major [T, N] (the kernel layout); the public fields are the reference's
logical [N, T] views (`.observations` is [N, T, d]) obtained by transposition without a copy.
For num_envs == 1 the views are [T] as in the reference (algorithm/ppo.py:355-358).
Flat indices follow the reference: i = n*T + t.
"""

from __future__ import annotations

from .. import _lib as L


class AbstractBuffer:
    pass


class RolloutBuffer(AbstractBuffer):
    def __init__(self, obs_tn=None, act_tn=None, rew_tn=None, done_tn=None, logp_tn=None, val_tn=None, ret_tn=None,
                 adv_tn=None, squeeze_env: bool = False, *, observations=None, actions=None, rewards=None, dones=None,
                 log_probs=None, values=None, returns=None, advantages=None, states=None, action_masks=None):
        """Positional arguments: the kernels' time-major [T][N] device tensors (what the training loop passes).
        Keyword arguments: the reference's constructor (buffer/rollout.py:21-50) with per-env [T] arrays or the
        vmapped [N, T] form; they are moved to the GPU and stored time-major."""
        if rewards is not None:
            torch = L.require_cuda()
            dev = torch.device("cuda", torch.cuda.current_device())
            rew = torch.as_tensor(rewards, dtype=torch.float32, device=dev)
            squeeze_env = rew.dim() == 1

            def tn(x, dtype):
                if x is None:
                    return None
                t = torch.as_tensor(x, device=dev).to(dtype)
                t = t.unsqueeze(0) if squeeze_env else t
                return t.transpose(0, 1).contiguous()  # [N, T, ...] -> [T, N, ...]

            act = torch.as_tensor(actions, device=dev)
            obs_tn = tn(observations, torch.float32)
            act_tn = tn(actions, torch.float32 if act.is_floating_point() else torch.int32)
            rew_tn, done_tn = tn(rewards, torch.float32), tn(dones, torch.uint8)
            logp_tn, val_tn = tn(log_probs, torch.float32), tn(values, torch.float32)
            ret_tn, adv_tn = tn(returns, torch.float32), tn(advantages, torch.float32)
        self._obs, self._act, self._rew, self._done = obs_tn, act_tn, rew_tn, done_tn
        self._logp, self._val, self._ret, self._adv = logp_tn, val_tn, ret_tn, adv_tn
        self._squeeze = squeeze_env
        self.states = states
        self.action_masks = action_masks

    # ---- reference-shaped views
    def _v(self, x):
        if x is None:
            return None
        y = x.transpose(0, 1)
        return y[0] if self._squeeze else y

    observations = property(lambda s: s._v(s._obs))
    actions = property(lambda s: s._v(s._act))
    rewards = property(lambda s: s._v(s._rew))
    dones = property(lambda s: s._v(s._done).bool())
    log_probs = property(lambda s: s._v(s._logp))
    values = property(lambda s: s._v(s._val))
    returns = property(lambda s: s._v(s._ret))
    advantages = property(lambda s: s._v(s._adv))

    @property
    def num_steps(self) -> int:
        return int(self._rew.shape[0])

    @property
    def num_envs(self) -> int:
        return int(self._rew.shape[1])

    @property
    def shape(self):
        return (self.num_steps,) if self._squeeze else (self.num_envs, self.num_steps)

    def compute_returns_and_advantages(self, last_value, gae_lambda, gamma, ev_sums=None) -> "RolloutBuffer":
        """buffer/rollout.py:64-79 (time-major single-pass kernel)."""
        torch = L.require_cuda()
        T, N = self.num_steps, self.num_envs
        adv = torch.empty_like(self._rew)
        ret = torch.empty_like(self._rew)
        lv = torch.as_tensor(last_value, dtype=torch.float32, device=self._rew.device).reshape(-1).expand(N).contiguous()
        L.check(L.lib.lrx_gae(L.ptr(self._rew), L.ptr(self._val), L.ptr(self._done), L.ptr(lv), L.ptr(adv),
                              L.ptr(ret), N, T, float(gamma), float(gae_lambda), L.LRX_LAYOUT_TN,
                              L.ptr(ev_sums), L.stream_ptr()), "lrx_gae")
        out = RolloutBuffer(self._obs, self._act, self._rew, self._done, self._logp, self._val, ret, adv,
                            self._squeeze)
        out.states, out.action_masks = self.states, self.action_masks
        return out

    def view(self, packed=None) -> L.LrxBatchView:
        """`packed`: optional [N*T, LRX_PACKED_FLOATS] float32 workspace; when given, the row-major copy of the buffer
        (lrx_ppo_pack_rows: one 64-byte record per row in reference order) is written and the view carries it, so the
        update's minibatch gather touches one line per row."""
        v = L.LrxBatchView(L.ptr(self._obs), L.ptr(self._act), L.ptr(self._logp), L.ptr(self._val),
                           L.ptr(self._ret), L.ptr(self._adv), self.num_envs, self.num_steps, None)
        if packed is not None:
            d = int(self._obs.shape[-1])
            L.check(L.lib.lrx_ppo_pack_rows(v, d, int(self._act.dtype.is_floating_point), L.ptr(packed), L.stream_ptr()),
                    "lrx_ppo_pack_rows")
            v.packed = L.ptr(packed)
        return v

    def batch_indices(self, batch_size: int, *, key=None):
        """base_buffer.py:64-88: permutation of N*T trimmed to a multiple of batch_size,
        reshaped [num_batches, batch_size]; int32 flat indices i = n*T + t."""
        torch = L.require_cuda()
        total = self.num_envs * self.num_steps
        dev = self._rew.device
        if key is None:
            idx = torch.arange(total, device=dev, dtype=torch.int32)
        else:
            idx = torch.empty(total, device=dev, dtype=torch.int32)
            L.check(L.lib.lrx_permutation(L.ptr(idx), total, int(key) & ((1 << 64) - 1), L.stream_ptr()),
                    "lrx_permutation")
        trim = total - (total % batch_size)
        return idx[:trim].reshape(-1, batch_size).contiguous()

    def flatten_axes(self):
        """base_buffer.py:38-62: dict of flat [N*T, ...] tensors in reference order (copies)."""
        out = {}
        for name in ("observations", "actions", "rewards", "dones", "log_probs", "values", "returns", "advantages"):
            x = getattr(self, name)
            if x is None:
                continue
            lead = 1 if self._squeeze else 2
            out[name] = x.reshape((-1,) + tuple(x.shape[lead:]))
        return out

    def gather(self, indices):
        """base_buffer.py:90-103."""
        flat = self.flatten_axes()
        idx = indices.long()
        return {k: v[idx] for k, v in flat.items()}


__all__ = ["AbstractBuffer", "RolloutBuffer"]
