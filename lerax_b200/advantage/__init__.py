"""Advantage estimators (reference: advantage/gae.py:36-66, bootstrapped.py:12-79, n_step.py:13-70).

`GAE(gamma, lam)(rewards, values, dones, last_value)` accepts the reference's per-env [T]
arrays or a batch [N, T] (the vmapped form) as CUDA tensors and runs the warp-scan kernel
(`lrx_gae`, LRX_LAYOUT_NT).  The training loop itself calls the time-major variant directly.
"""

from __future__ import annotations

from .. import _lib as L


class AbstractAdvantageEstimator:
    pass


def _prepare(rewards, values, dones, last_value, what):
    """[T] or [N, T] inputs -> contiguous [N, T] CUDA tensors (LRX_LAYOUT_NT) + empty outputs."""
    torch = L.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    rewards = torch.as_tensor(rewards, dtype=torch.float32, device=dev)
    single = rewards.dim() == 1
    r = rewards.reshape(1, -1) if single else rewards
    if r.dim() != 2:
        raise ValueError(f"{what} expects [T] or [N, T] arrays")
    n, t = r.shape
    r = r.contiguous()
    v = torch.as_tensor(values, dtype=torch.float32, device=dev).reshape(n, t).contiguous()
    d = torch.as_tensor(dones, device=dev).to(torch.uint8).reshape(n, t).contiguous()
    lv = torch.as_tensor(last_value, dtype=torch.float32, device=dev).reshape(-1).expand(n).contiguous()
    return single, n, t, r, v, d, lv, torch.empty_like(r), torch.empty_like(r)


class GAE(AbstractAdvantageEstimator):
    def __init__(self, gamma: float = 0.99, lam: float = 0.95):
        self.gamma = gamma
        self.lam = lam

    def __call__(self, rewards, values, dones, last_value):
        single, n, t, r, v, d, lv, adv, ret = _prepare(rewards, values, dones, last_value, "GAE")
        L.check(L.lib.lrx_gae(L.ptr(r), L.ptr(v), L.ptr(d), L.ptr(lv), L.ptr(adv), L.ptr(ret), n, t,
                              float(self.gamma), float(self.lam), L.LRX_LAYOUT_NT, None, L.stream_ptr()), "lrx_gae")
        if single:
            return adv[0], ret[0]
        return adv, ret


class BootstrappedReturn(AbstractAdvantageEstimator):
    """Discounted returns bootstrapped at the segment boundary (advantage/bootstrapped.py:52-79)."""

    def __init__(self, gamma: float = 0.99):
        self.gamma = gamma

    def __call__(self, rewards, values, dones, last_value):
        single, n, t, r, v, d, lv, adv, ret = _prepare(rewards, values, dones, last_value, "BootstrappedReturn")
        L.check(L.lib.lrx_returns(L.ptr(r), L.ptr(v), L.ptr(d), L.ptr(lv), L.ptr(adv), L.ptr(ret), n, t,
                                  float(self.gamma), 0, L.LRX_LAYOUT_NT, L.stream_ptr()), "lrx_returns")
        if single:
            return adv[0], ret[0]
        return adv, ret


def discounted_returns(rewards, dones, final_value, gamma):
    """advantage/bootstrapped.py:12-49."""
    torch = L.require_cuda()
    zeros = torch.zeros_like(torch.as_tensor(rewards, dtype=torch.float32))
    return BootstrappedReturn(gamma)(rewards, zeros, dones, final_value)[1]


class NStepReturn(AbstractAdvantageEstimator):
    """N-step bootstrapped returns (advantage/n_step.py:13-70)."""

    def __init__(self, n: int, gamma: float = 0.99):
        if n < 1:
            raise ValueError(f"n must be at least 1, got {n}")
        self.n = n
        self.gamma = gamma

    def __call__(self, rewards, values, dones, last_value):
        single, n, t, r, v, d, lv, adv, ret = _prepare(rewards, values, dones, last_value, "NStepReturn")
        L.check(L.lib.lrx_returns(L.ptr(r), L.ptr(v), L.ptr(d), L.ptr(lv), L.ptr(adv), L.ptr(ret), n, t,
                                  float(self.gamma), int(self.n), L.LRX_LAYOUT_NT, L.stream_ptr()), "lrx_returns")
        if single:
            return adv[0], ret[0]
        return adv, ret


__all__ = ["AbstractAdvantageEstimator", "BootstrappedReturn", "GAE", "NStepReturn", "discounted_returns"]
