"""Stand-in for jax 0.9.0.1 (uv.lock:848-849) on torch CPU tensors -- TEST INFRASTRUCTURE, see
oracle/_refexec/__init__.py.  fp32 / int32 defaults (x64 disabled), jnp semantics where they differ from torch's
(clip = minimum(maximum(.)), population std/var, remainder with the divisor's sign)."""

import torch

__version__ = "0.9.0.1-refexec-shim"
Array = torch.Tensor

from . import debug, lax, nn, numpy, random, tree, typing  # noqa: E402,F401
from . import tree_util  # noqa: E402,F401
from ._vmap import vmap  # noqa: E402,F401


def jit(fn=None, **_):
    return fn if fn is not None else (lambda f: f)
