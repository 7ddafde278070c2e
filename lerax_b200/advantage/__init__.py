"""GAE advantage estimator (reference: advantage/gae.py:36-66).

`GAE(gamma, lam)(rewards, values, dones, last_value)` accepts the reference's per-env [T]
arrays or a batch [N, T] (the vmapped form) as CUDA tensors and runs the warp-scan kernel
(`lrx_gae`, LRX_LAYOUT_NT).  The training loop itself calls the time-major variant directly.
"""

from __future__ import annotations

from .. import _lib as L


class AbstractAdvantageEstimator:
    pass


class GAE(AbstractAdvantageEstimator):
    def __init__(self, gamma: float = 0.99, lam: float = 0.95):
        self.gamma = gamma
        self.lam = lam

    def __call__(self, rewards, values, dones, last_value):
        torch = L.require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device())
        rewards = torch.as_tensor(rewards, dtype=torch.float32, device=dev)
        single = rewards.dim() == 1
        r = rewards.reshape(1, -1) if single else rewards
        if r.dim() != 2:
            raise ValueError("GAE expects [T] or [N, T] arrays")
        n, t = r.shape
        r = r.contiguous()
        v = torch.as_tensor(values, dtype=torch.float32, device=dev).reshape(n, t).contiguous()
        d = torch.as_tensor(dones, device=dev).to(torch.uint8).reshape(n, t).contiguous()
        lv = torch.as_tensor(last_value, dtype=torch.float32, device=dev).reshape(n).contiguous()
        adv = torch.empty_like(r)
        ret = torch.empty_like(r)
        L.check(L.lib.lrx_gae(L.ptr(r), L.ptr(v), L.ptr(d), L.ptr(lv), L.ptr(adv), L.ptr(ret), n, t,
                              float(self.gamma), float(self.lam), L.LRX_LAYOUT_NT, None, L.stream_ptr()), "lrx_gae")
        if single:
            return adv[0], ret[0]
        return adv, ret


__all__ = ["AbstractAdvantageEstimator", "GAE"]
