"""MLPActorCriticPolicy with the reference's constructor and method surface.

Reference: policy/actor_critic/mlp.py:64-178, policy/actor.py:40-88,194-261,
policy/actor_critic/base_actor_critic.py:37-87.  Parameters live in ONE flat float32 CUDA
vector in Equinox leaf order (encoder.layers[i].{weight,bias}, value_head.layers[i]...,
action_head.mlp.layers[i]..., action_head.action_dist.mapping.{weight,bias}, log_std) --
the layout the kernels and the Adam state share; `.encoder.layers[i].weight` etc. are views.
"""

from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np

from .. import _lib as L
from ..random import Key, split
from ..space import Box, Discrete


class AbstractPolicy:
    pass


class AbstractActorCriticPolicy(AbstractPolicy):
    pass


def _relu_marker(x):  # placeholder identity for the `activation=` default
    raise NotImplementedError


class MLPActorCriticPolicy(AbstractActorCriticPolicy):
    name = "MLPActorCriticPolicy"

    def __init__(self, env, *, feature_size: int = 16, feature_width: int = 64, feature_depth: int = 2,
                 value_width: int = 64, value_depth: int = 2, action_width: int = 64, action_depth: int = 2,
                 activation=None, log_std_init: float = 0.0, key: Key | int, device=None, _params=None):
        if activation is not None and getattr(activation, "__name__", "") != "relu":
            raise NotImplementedError("lerax_b200 kernels implement the reference default activation (relu) only")
        self.action_space = env.action_space
        self.observation_space = env.observation_space
        obs_dim = int(self.observation_space.flat_size)
        if isinstance(self.action_space, Discrete):
            n_out, is_box = int(self.action_space.n), False
        elif isinstance(self.action_space, Box) and self.action_space.shape == ():
            n_out, is_box = 1, True
        else:
            raise NotImplementedError(f"action space {self.action_space!r} is outside the classic-control hot path")
        if n_out > L.LRX_MAX_NOUT:
            raise NotImplementedError(f"at most {L.LRX_MAX_NOUT} discrete actions are supported")
        ad = max(0, action_depth - 1)  # actor.py:232
        self._dims = (
            (obs_dim,) + (feature_width,) * feature_depth + (feature_size,),
            (feature_size,) + (value_width,) * value_depth + (1,),
            (feature_size,) + (action_width,) * ad + (feature_size,) + (n_out,),
        )
        self._relu = ((1,) * feature_depth + (0,), (1,) * value_depth + (0,), (1,) * ad + (0, 0))
        for d in self._dims:
            if len(d) - 1 > L.LRX_MAX_LAYERS:
                raise NotImplementedError(f"at most {L.LRX_MAX_LAYERS} Linear layers per chain are supported")
        self.obs_dim, self.n_out, self.is_box = obs_dim, n_out, is_box
        self._layout = []  # (chain, layer, w_off, b_off, in, out)
        off = 0
        for c, chain in enumerate(self._dims):
            for i in range(len(chain) - 1):
                k, n = chain[i], chain[i + 1]
                self._layout.append((c, i, off, off + k * n, k, n))
                off += k * n + n
        self.log_std_offset = off if is_box else -1
        self.n_params = off + (1 if is_box else 0)

        if _params is not None:
            self.params = _params
        else:
            # eqx.nn.Linear init: weight, bias ~ U(-1/sqrt(in), 1/sqrt(in)); host-side, one H2D copy
            rng = np.random.Generator(np.random.PCG64(int(key) & ((1 << 64) - 1)))
            p = np.zeros(self.n_params, np.float32)
            for (_, _, w, b, k, n) in self._layout:
                lim = 1.0 / math.sqrt(k)
                p[w:w + k * n] = rng.uniform(-lim, lim, k * n)
                p[b:b + n] = rng.uniform(-lim, lim, n)
            if is_box:
                p[self.log_std_offset] = log_std_init
            torch = L.require_cuda() if device is None or str(device) != "cpu" else __import__("torch")
            self.params = torch.from_numpy(p).to(device if device is not None else "cuda")

    # ------------------------------------------------------------------ structure
    def desc(self) -> L.LrxPolicyDesc:
        d = L.LrxPolicyDesc()
        d.obs_dim, d.n_out, d.is_box, d.n_params = self.obs_dim, self.n_out, int(self.is_box), self.n_params
        for c in range(3):
            d.n_layers[c] = len(self._dims[c]) - 1
            for i, w in enumerate(self._dims[c]):
                d.dims[c][i] = w
            for i, r in enumerate(self._relu[c]):
                d.relu[c][i] = r
        return d

    def is_default_widths(self) -> bool:
        """mlp.py:64-78 defaults (encoder d->64->64->16, value head 16->64->64->1, action head 16->64->16->n_out, d <= 8):
        the shapes the specialised tensor-core kernels serve."""
        d, n = self.obs_dim, self.n_out
        return d <= 8 and tuple(self._dims) == ((d, 64, 64, 16), (16, 64, 64, 1), (16, 64, 16, n))

    def with_params(self, params) -> "MLPActorCriticPolicy":
        new = object.__new__(MLPActorCriticPolicy)
        new.__dict__.update(self.__dict__)
        new.params = params
        return new

    def to(self, device) -> "MLPActorCriticPolicy":
        return self.with_params(self.params.to(device))

    def numpy(self) -> np.ndarray:
        return self.params.detach().cpu().numpy()

    @classmethod
    def from_numpy(cls, env, params: np.ndarray, device="cuda", **kw):
        import torch

        pol = cls(env, key=0, device="cpu", **kw)
        if params.shape != (pol.n_params,):
            raise ValueError(f"expected {pol.n_params} parameters, got {params.shape}")
        return pol.with_params(torch.from_numpy(np.ascontiguousarray(params, np.float32)).to(device))

    def _chain_views(self, c):
        layers = []
        for (cc, i, w, b, k, n) in self._layout:
            if cc == c:
                layers.append(SimpleNamespace(weight=self.params[w:w + k * n].view(n, k), bias=self.params[b:b + n],
                                              in_features=k, out_features=n))
        return layers

    @property
    def encoder(self):
        return SimpleNamespace(layers=tuple(self._chain_views(0)))

    @property
    def value_head(self):
        return SimpleNamespace(layers=tuple(self._chain_views(1)))

    @property
    def action_head(self):
        ls = self._chain_views(2)
        dist = SimpleNamespace(mapping=ls[-1])
        if self.is_box:
            dist.log_std = self.params[self.log_std_offset]
        return SimpleNamespace(mlp=SimpleNamespace(layers=tuple(ls[:-1])), action_dist=dist)

    @property
    def log_std(self):
        return self.params[self.log_std_offset] if self.is_box else None

    # ------------------------------------------------------------------ interface
    def reset(self, *, key=None):
        return None  # stateless (mlp.py:112-113)

    def _forward(self, observation):
        torch = L.require_cuda()
        if not self.params.is_cuda:
            raise L.LeraxB200Error("policy parameters are on the host; call policy.to('cuda') first")
        obs = observation.to(device=self.params.device, dtype=torch.float32)
        single = obs.dim() == 1
        obs = obs.reshape(-1, self.obs_dim).contiguous()
        B = obs.shape[0]
        value = torch.empty((B,), dtype=torch.float32, device=obs.device)
        head = torch.empty((B, self.n_out), dtype=torch.float32, device=obs.device)
        d = self.desc()
        L.check(L.lib.lrx_policy_forward(d, L.ptr(self.params), L.ptr(obs), L.ptr(value), L.ptr(head), B,
                                         L.stream_ptr()), "lrx_policy_forward")
        return value, head, single

    def value(self, state, observation):
        """mlp.py:154-157."""
        value, _, single = self._forward(observation)
        return None, (value[0] if single else value)

    def logits_or_loc(self, observation):
        _, head, single = self._forward(observation)
        return head[0] if single else head

    def __call__(self, state, observation, *, key: Key | int | None = None, action_mask=None):
        """mlp.py:115-131: mode() when key is None, else a sample."""
        torch = L.require_cuda()
        if action_mask is not None:
            raise NotImplementedError("action masks are outside the classic-control hot path")
        value, head, single = self._forward(observation)
        if self.is_box:
            loc = head[:, 0]
            if key is None:
                act = loc
            else:
                g = torch.Generator(device=loc.device).manual_seed(int(key) & ((1 << 63) - 1))
                act = loc + torch.exp(self.params[self.log_std_offset]) * torch.randn(loc.shape, generator=g, device=loc.device)
        else:
            if key is None:
                act = torch.argmax(head, dim=-1).to(torch.int32)
            else:
                g = torch.Generator(device=head.device).manual_seed(int(key) & ((1 << 63) - 1))
                u = torch.rand(head.shape, generator=g, device=head.device).clamp_(1e-20, 1.0)
                act = torch.argmax(torch.log_softmax(head, -1) - torch.log(-torch.log(u)), dim=-1).to(torch.int32)
        return None, (act[0] if single else act)


__all__ = ["AbstractPolicy", "AbstractActorCriticPolicy", "MLPActorCriticPolicy"]
