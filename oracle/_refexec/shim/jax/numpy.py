"""jax.numpy on torch CPU tensors: the calls the reference's on-policy path makes.  Defaults follow JAX with x64
disabled (Python floats -> float32, ints -> int32); semantic differences from torch are restated:
clip(x, lo, hi) = minimum(maximum(x, lo), hi) (also for the sub-gradients), std / var are population moments,
`%` / remainder take the divisor's sign (torch.remainder already does)."""

from __future__ import annotations

import builtins
import math

import numpy as np
import torch

pi, inf, nan, e = math.pi, math.inf, math.nan, math.e
newaxis = None
ndarray = torch.Tensor
float32, float64, int32, int64, uint8, bool_ = torch.float32, torch.float64, torch.int32, torch.int64, torch.uint8, torch.bool
float_, int_ = torch.float32, torch.int32


def _dtype(dt):
    if dt is None:
        return None
    if dt is float or dt is builtins.float:
        return torch.float32
    if dt is int:
        return torch.int32
    if dt is bool:
        return torch.bool
    if isinstance(dt, torch.dtype):
        return dt
    return {np.dtype("float32"): torch.float32, np.dtype("float64"): torch.float32, np.dtype("int32"): torch.int32,
            np.dtype("int64"): torch.int32, np.dtype("bool"): torch.bool, np.dtype("uint8"): torch.uint8}[np.dtype(dt)]


def _default_dtype(x):
    if isinstance(x, bool):
        return torch.bool
    if isinstance(x, int):
        return torch.int32
    if isinstance(x, float):
        return torch.float32
    return None


def asarray(x, dtype=None):
    dt = _dtype(dtype)
    if isinstance(x, torch.Tensor):
        return x if dt is None or x.dtype == dt else x.to(dt)
    if isinstance(x, (np.ndarray, np.generic)):
        t = torch.from_numpy(np.asarray(x).copy())
        if dt is None:
            dt = {torch.float64: torch.float32, torch.int64: torch.int32}.get(t.dtype, t.dtype)
        return t.to(dt)
    if isinstance(x, (list, tuple)):
        if len(x) == 0:
            return torch.zeros((0,), dtype=dt or torch.float32)
        parts = [asarray(v) for v in x]
        if builtins.any(p.dtype.is_floating_point for p in parts):
            parts = [p.to(torch.float32) for p in parts]
        t = torch.stack(parts)
        return t if dt is None else t.to(dt)
    return torch.tensor(x, dtype=dt or _default_dtype(x))


def array(x, dtype=None, copy=True):
    t = asarray(x, dtype)
    return t.clone() if (isinstance(x, torch.Tensor) and t is x) else t


def _t(x, like=None):
    if isinstance(x, torch.Tensor):
        return x
    if like is not None and isinstance(x, (int, float)) and not isinstance(x, bool):
        return torch.tensor(x, dtype=like.dtype if (like.dtype.is_floating_point or isinstance(x, int)) else torch.float32)
    return asarray(x)


def _patch_tensor():
    torch.Tensor.astype = lambda self, dt: self.to(_dtype(dt))
    if not hasattr(torch.Tensor, "_refexec_squeeze"):
        torch.Tensor._refexec_squeeze = True


_patch_tensor()


class _Finfo:
    def __init__(self, dt):
        f = torch.finfo(_dtype(dt) if not isinstance(dt, torch.dtype) else dt)
        self.eps, self.max, self.min, self.tiny = f.eps, f.max, f.min, f.tiny


def finfo(dt):
    return _Finfo(dt)


def sin(x): return torch.sin(_t(x))
def cos(x): return torch.cos(_t(x))
def exp(x): return torch.exp(_t(x))
def log(x): return torch.log(_t(x))
def sqrt(x): return torch.sqrt(_t(x))
def abs(x): return torch.abs(_t(x))  # noqa: A001
def square(x): x = _t(x); return x * x
def floor(x): return torch.floor(_t(x))
def tanh(x): return torch.tanh(_t(x))
def sign(x): return torch.sign(_t(x))
def isfinite(x): return torch.isfinite(_t(x))
def isnan(x): return torch.isnan(_t(x))
def arctan2(a, b): a = _t(a); return torch.atan2(a, _t(b, a))
def logical_not(x): return ~_t(x)
def logical_and(a, b): return _t(a) & _t(b)
def logical_or(a, b): return _t(a) | _t(b)


def minimum(a, b):
    a = _t(a, b if isinstance(b, torch.Tensor) else None)
    return torch.minimum(a, _t(b, a).to(a.dtype) if not isinstance(b, torch.Tensor) else b)


def maximum(a, b):
    a = _t(a, b if isinstance(b, torch.Tensor) else None)
    return torch.maximum(a, _t(b, a).to(a.dtype) if not isinstance(b, torch.Tensor) else b)


def clip(x, min=None, max=None, a_min=None, a_max=None):  # noqa: A002
    lo = min if min is not None else a_min
    hi = max if max is not None else a_max
    x = _t(x)
    if lo is not None:
        x = maximum(x, lo)
    if hi is not None:
        x = minimum(x, hi)
    return x


def remainder(a, b):
    a = _t(a)
    return torch.remainder(a, _t(b, a))


mod = remainder


def where(c, a=None, b=None):
    c = _t(c)
    if isinstance(a, torch.Tensor):
        b = _t(b, a)
    elif isinstance(b, torch.Tensor):
        a = _t(a, b)
    else:
        a, b = asarray(a), asarray(b)
        if a.dtype != b.dtype:
            a, b = a.to(torch.float32), b.to(torch.float32)
    return torch.where(c, a, b)


def _axis_kw(axis):
    return {} if axis is None else {"dim": axis}


def mean(x, axis=None, keepdims=False):
    x = _t(x)
    return x.mean() if axis is None else x.mean(dim=axis, keepdim=keepdims)


def sum(x, axis=None, keepdims=False):  # noqa: A001
    x = _t(x)
    return x.sum() if axis is None else x.sum(dim=axis, keepdim=keepdims)


def var(x, axis=None, ddof=0):
    x = _t(x)
    return torch.var(x, correction=ddof) if axis is None else torch.var(x, dim=axis, correction=ddof)


def std(x, axis=None, ddof=0):
    return torch.sqrt(var(x, axis, ddof))


def all(x, axis=None):  # noqa: A001
    x = _t(x)
    return x.all() if axis is None else x.all(dim=axis)


def any(x, axis=None):  # noqa: A001
    x = _t(x)
    return x.any() if axis is None else x.any(dim=axis)


def max(x, axis=None):  # noqa: A001
    x = _t(x)
    return x.max() if axis is None else x.max(dim=axis).values


def min(x, axis=None):  # noqa: A001
    x = _t(x)
    return x.min() if axis is None else x.min(dim=axis).values


def argmax(x, axis=None):
    x = _t(x)
    return (x.argmax() if axis is None else x.argmax(dim=axis)).to(torch.int32)


def cumsum(x, axis=0):
    return torch.cumsum(_t(x), dim=axis)


def stack(xs, axis=0):
    return torch.stack([asarray(x) for x in xs], dim=axis)


def concatenate(xs, axis=0):
    parts = [asarray(x) for x in xs]
    parts = [p.reshape(1) if p.ndim == 0 else p for p in parts]
    return torch.cat(parts, dim=axis)


def unstack(x, axis=0):
    return list(torch.unbind(_t(x), dim=axis))


def split(x, n, axis=0):
    x = _t(x)
    if isinstance(n, int):
        return list(torch.chunk(x, n, dim=axis))
    idx = [0] + [int(i) for i in n] + [x.shape[axis]]
    return [x.narrow(axis, a, b - a) for a, b in zip(idx[:-1], idx[1:])]


def take(x, indices, axis=None):
    x = _t(x)
    idx = _t(indices).to(torch.int64)
    if axis is None:
        return x.reshape(-1)[idx]
    return torch.index_select(x, axis, idx.reshape(-1)).reshape(x.shape[:axis] + idx.shape + x.shape[axis + 1:])


def moveaxis(x, src, dst):
    return torch.movedim(_t(x), src, dst)


def broadcast_to(x, shape):
    return torch.broadcast_to(_t(x), tuple(shape))


def broadcast_arrays(*xs):
    return list(torch.broadcast_tensors(*[_t(x) for x in xs]))


def reshape(x, shape):
    return _t(x).reshape(shape)


def ravel(x):
    return _t(x).reshape(-1)


def squeeze(x, axis=None):
    return _t(x).squeeze() if axis is None else _t(x).squeeze(axis)


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def zeros(shape, dtype=None): return torch.zeros(shape if not isinstance(shape, int) else (shape,), dtype=_dtype(dtype) or torch.float32)
def ones(shape, dtype=None): return torch.ones(shape if not isinstance(shape, int) else (shape,), dtype=_dtype(dtype) or torch.float32)
def empty(shape, dtype=None): return zeros(shape, dtype)
def full(shape, v, dtype=None):
    if isinstance(shape, int):
        shape = (shape,)
    v = _t(v)
    return torch.full(tuple(shape), 0, dtype=_dtype(dtype) or v.dtype) + v.to(_dtype(dtype) or v.dtype)
def zeros_like(x, dtype=None): return torch.zeros_like(_t(x), dtype=_dtype(dtype))
def ones_like(x, dtype=None): return torch.ones_like(_t(x), dtype=_dtype(dtype))
def full_like(x, v, dtype=None): return torch.full_like(_t(x), v, dtype=_dtype(dtype))
def arange(*a, dtype=None): return torch.arange(*a, dtype=_dtype(dtype) or (torch.int32 if builtins.all(isinstance(v, int) for v in a) else torch.float32))
def linspace(a, b, n, dtype=None): return torch.linspace(float(a), float(b), int(n), dtype=_dtype(dtype) or torch.float32)


def array_equal(a, b):
    a, b = _t(a), _t(b)
    return torch.tensor(a.shape == b.shape and bool((a == b).all()))


def isscalar(x):
    return isinstance(x, (int, float, bool)) or (isinstance(x, torch.Tensor) and x.ndim == 0)


def result_type(*xs):
    return torch.result_type(*[_t(x) for x in xs[:2]])
