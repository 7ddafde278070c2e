"""jax.random with INJECTED noise.  RNG streams are not part of the parity contract: a key is a 0-d int64
tensor naming a node of the key tree; `split` derives children `path + (j,)`; every sampler asks the installed
handler for the raw draw belonging to its key's path, and applies jax's own post-processing (uniform: u * (max-min)
+ min, clamped at min; categorical: argmax(logits + Gumbel); normal: the N(0,1) draw itself)."""

from __future__ import annotations

import torch

_paths: dict[int, tuple] = {}
_next_id = [1]
_handler = [None]


def reset_registry():
    _paths.clear()
    _next_id[0] = 1


def set_noise_handler(fn):
    """fn(path: tuple, kind: str, shape: tuple) -> tensor / array of raw draws."""
    _handler[0] = fn


def make_key(path) -> torch.Tensor:
    i = _next_id[0]
    _next_id[0] += 1
    _paths[i] = tuple(path)
    return torch.tensor(i, dtype=torch.int64)


def key_path(key) -> tuple:
    return _paths[int(key)]


def key(seed):
    return make_key(("seed", int(seed)))


PRNGKey = key


def split(key, num=2):
    p = key_path(key)
    if int(num) == 0:
        return torch.zeros((0,), dtype=torch.int64)
    return torch.stack([make_key(p + (j,)) for j in range(int(num))])


def _draw(key, kind, shape):
    if _handler[0] is None:
        raise RuntimeError("jax.random stand-in: no noise handler installed")
    from . import numpy as jnp

    out = jnp.asarray(_handler[0](key_path(key), kind, tuple(shape)))
    return out


def uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0):
    from . import numpy as jnp

    u = _draw(key, "uniform", shape).to(torch.float32)
    lo, hi = jnp.asarray(minval).to(torch.float32), jnp.asarray(maxval).to(torch.float32)
    return torch.maximum(lo, u * (hi - lo) + lo)  # jax._src.random._uniform


def normal(key, shape=(), dtype=None):
    return _draw(key, "normal", shape).to(torch.float32)


def gumbel(key, shape=(), dtype=None):
    return _draw(key, "gumbel", shape).to(torch.float32)


def categorical(key, logits, axis=-1, shape=None):
    g = gumbel(key, logits.shape)
    return torch.argmax(logits + g, dim=axis).to(torch.int32)


def permutation(key, x):
    n = int(x) if not isinstance(x, torch.Tensor) or x.ndim == 0 else x.shape[0]
    return _draw(key, "permutation", (n,)).to(torch.int32)


def _unsupported(name):
    def f(*a, **k):
        raise NotImplementedError(f"jax.random.{name} is not part of the on-policy path")

    return f


exponential, choice, randint, bernoulli = (_unsupported(n) for n in ("exponential", "choice", "randint", "bernoulli"))
