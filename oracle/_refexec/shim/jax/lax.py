"""lax.scan / lax.cond as plain Python control flow over pytrees (utils.py:65-136 `filter_scan` and
algorithm/ppo.py:248-263 are the call sites)."""

import equinox as eqx
import torch


def scan(f, init, xs=None, length=None, reverse=False, unroll=1, _split_transpose=False):
    if xs is None:
        n = int(length)
    else:
        sizes = {leaf.shape[0] for leaf in eqx.tree_leaves(xs) if isinstance(leaf, torch.Tensor)}
        if len(sizes) != 1:
            raise ValueError(f"scan: inconsistent leading sizes {sizes}")
        n = sizes.pop()
    order = range(n - 1, -1, -1) if reverse else range(n)
    carry, ys = init, [None] * n
    for i in order:
        x = None if xs is None else eqx.tree_map(lambda leaf: leaf[i] if isinstance(leaf, torch.Tensor) else leaf, xs)
        carry, y = f(carry, x)
        ys[i] = y
    if n == 0:
        return carry, None
    return carry, eqx.tree_map(eqx._stack_leaves, ys[0], *ys[1:])


def cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def select(pred, a, b):
    return torch.where(pred, a, b)


def stop_gradient(x):
    return eqx.tree_map(lambda t: t.detach() if isinstance(t, torch.Tensor) else t, x)


def dynamic_slice_in_dim(x, start, size, axis=0):
    return torch.narrow(x, axis, int(start), size)
