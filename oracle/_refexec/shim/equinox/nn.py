"""eqx.nn.Linear / eqx.nn.MLP of equinox 0.13.6: y = W[out,in] x + b, "scalar" sizes squeeze,
activation on hidden layers only, identity final activation (policy/actor_critic/mlp.py:84-100,
policy/actor.py:56-62,85,236-243 are the call sites).  Initial values are U(+-1/sqrt(in)) from a NumPy
generator: parity runs overwrite them with injected parameters."""

from __future__ import annotations

import math
from typing import Callable, Literal, Union

import numpy as np
import torch

from . import Module, field

_INIT_RNG = np.random.default_rng(0)


def _identity(x):
    return x


def _relu(x):
    return torch.relu(x)


class Linear(Module):
    weight: torch.Tensor
    bias: torch.Tensor
    in_features: Union[int, Literal["scalar"]] = field(static=True)
    out_features: Union[int, Literal["scalar"]] = field(static=True)
    use_bias: bool = field(static=True)

    def __init__(self, in_features, out_features, use_bias=True, dtype=None, *, key=None):
        i = 1 if in_features == "scalar" else int(in_features)
        o = 1 if out_features == "scalar" else int(out_features)
        lim = 1.0 / math.sqrt(i)
        self.weight = torch.from_numpy(_INIT_RNG.uniform(-lim, lim, (o, i)).astype(np.float32))
        self.bias = torch.from_numpy(_INIT_RNG.uniform(-lim, lim, (o,)).astype(np.float32)) if use_bias else None
        self.in_features, self.out_features, self.use_bias = in_features, out_features, use_bias

    def __call__(self, x, *, key=None):
        if self.in_features == "scalar":
            x = x.reshape(1)
        y = self.weight @ x
        if self.bias is not None:
            y = y + self.bias
        if self.out_features == "scalar":
            y = y.squeeze(0)
        return y


class MLP(Module):
    layers: tuple
    activation: Callable = field(static=True)
    final_activation: Callable = field(static=True)
    use_bias: bool = field(static=True)
    use_final_bias: bool = field(static=True)
    in_size: Union[int, Literal["scalar"]] = field(static=True)
    out_size: Union[int, Literal["scalar"]] = field(static=True)
    width_size: int = field(static=True)
    depth: int = field(static=True)

    def __init__(self, in_size, out_size, width_size, depth, activation=_relu, final_activation=_identity,
                 use_bias=True, use_final_bias=True, dtype=None, *, key=None):
        layers = []
        if depth == 0:
            layers.append(Linear(in_size, out_size, use_final_bias))
        else:
            layers.append(Linear(in_size, width_size, use_bias))
            for _ in range(depth - 1):
                layers.append(Linear(width_size, width_size, use_bias))
            layers.append(Linear(width_size, out_size, use_final_bias))
        self.layers = tuple(layers)
        self.activation, self.final_activation = activation, final_activation
        self.use_bias, self.use_final_bias = use_bias, use_final_bias
        self.in_size, self.out_size, self.width_size, self.depth = in_size, out_size, width_size, depth

    def __call__(self, x, *, key=None):
        for layer in self.layers[:-1]:
            x = self.activation(layer(x))
        return self.final_activation(self.layers[-1](x))
