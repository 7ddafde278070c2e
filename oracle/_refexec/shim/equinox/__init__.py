"""Stand-in for equinox 0.13.6 (uv.lock:420-422) -- TEST INFRASTRUCTURE, see oracle/_refexec/__init__.py.

Only what the reference's on-policy path touches: `Module` (fields from annotations, custom or
generated __init__), pytree flatten / map, `partition` / `combine` / `filter`, `tree_at`,
`filter_vmap` (a Python loop + stack), `filter_value_and_grad` (torch autograd), `apply_updates`,
`error_if`, and `nn.Linear` / `nn.MLP` (y = W x + b, activation on hidden layers only).
Arrays are torch CPU tensors.
"""

from __future__ import annotations

import abc
import dataclasses
import typing

import torch

_MISSING = dataclasses.MISSING


# --------------------------------------------------------------------------- field markers
class _AbstractMarker:
    def __init__(self, kind):
        self.kind = kind

    def __repr__(self):
        return f"{self.kind}[...]"


class AbstractVar:
    def __class_getitem__(cls, item):
        return _AbstractMarker("AbstractVar")


class AbstractClassVar:
    def __class_getitem__(cls, item):
        return _AbstractMarker("AbstractClassVar")


class _FieldSpec:
    def __init__(self, static=False, converter=None, default=_MISSING, default_factory=_MISSING):
        self.static, self.converter, self.default, self.default_factory = static, converter, default, default_factory


def field(*, static=False, converter=None, default=_MISSING, default_factory=_MISSING, **_):
    return _FieldSpec(static, converter, default, default_factory)


def _annotation_kind(ann) -> str:
    s = ann if isinstance(ann, str) else repr(ann)
    if isinstance(ann, _AbstractMarker):
        return ann.kind
    if "AbstractClassVar" in s or "ClassVar" in s:
        return "ClassVar"
    if "AbstractVar" in s:
        return "AbstractVar"
    return "field"


class _ModuleMeta(abc.ABCMeta):
    def __new__(mcs, name, bases, ns, **kw):
        cls = super().__new__(mcs, name, bases, ns, **kw)
        fields, specs = [], {}
        for base in reversed(cls.__mro__):
            ann = base.__dict__.get("__annotations__", {})
            for k, a in ann.items():
                kind = _annotation_kind(a)
                if kind == "ClassVar":
                    if k in fields:
                        fields.remove(k)
                    continue
                if kind == "AbstractVar":
                    continue
                if k not in fields:
                    fields.append(k)
                v = base.__dict__.get(k, _MISSING)
                if isinstance(v, _FieldSpec):
                    specs[k] = v
                elif v is not _MISSING and not isinstance(v, property):
                    specs[k] = _FieldSpec(default=v)
        fields = [f for f in fields if not isinstance(getattr(cls, f, None), property)]
        cls.__eqx_fields__ = tuple(fields)
        cls.__eqx_specs__ = specs
        for k, v in specs.items():  # dataclass-style: the class attribute becomes the default value
            if isinstance(cls.__dict__.get(k), _FieldSpec):
                if v.default is not _MISSING:
                    setattr(cls, k, v.default)
                else:
                    delattr(cls, k)
        return cls

    def __call__(cls, *args, **kwargs):
        self = cls.__new__(cls)
        has_init = any("__init__" in b.__dict__ for b in cls.__mro__ if b not in (Module, object, typing.Generic))
        if has_init:
            cls.__init__(self, *args, **kwargs)
        else:
            names = list(cls.__eqx_fields__)
            if len(args) > len(names):
                raise TypeError(f"{cls.__name__}: too many positional arguments")
            vals = dict(zip(names, args))
            for k, v in kwargs.items():
                if k not in names or k in vals:
                    raise TypeError(f"{cls.__name__}: unexpected argument {k}")
                vals[k] = v
            for k in names:
                if k not in vals:
                    sp = cls.__eqx_specs__.get(k)
                    if sp is not None and sp.default is not _MISSING:
                        vals[k] = sp.default
                    elif sp is not None and sp.default_factory is not _MISSING:
                        vals[k] = sp.default_factory()
                    else:
                        raise TypeError(f"{cls.__name__}: missing argument {k}")
                object.__setattr__(self, k, vals[k])
        for k, sp in cls.__eqx_specs__.items():
            if sp.converter is not None and hasattr(self, k):
                object.__setattr__(self, k, sp.converter(getattr(self, k)))
        post = getattr(self, "__check_init__", None)
        if post is not None:
            post()
        return self


class Module(metaclass=_ModuleMeta):
    def __repr__(self):
        inner = ", ".join(f"{k}={getattr(self, k, None)!r}" for k in type(self).__eqx_fields__)
        return f"{type(self).__name__}({inner})"


def _static_fields(cls):
    return {k for k, sp in cls.__eqx_specs__.items() if sp.static}


# ----------------------------------------------------------------------------- pytrees
def is_array(x) -> bool:
    return isinstance(x, torch.Tensor)


def is_inexact_array(x) -> bool:
    return isinstance(x, torch.Tensor) and (x.dtype.is_floating_point or x.dtype.is_complex)


def is_array_like(x) -> bool:
    import numpy as np

    return isinstance(x, (torch.Tensor, np.ndarray, np.generic, int, float, complex, bool))


def if_array(axis):
    return lambda x: axis if is_array(x) else None


def _is_namedtuple(x):
    return isinstance(x, tuple) and hasattr(type(x), "_fields")


def _children(node):
    """(kind, keys, children, rebuild) of an internal node, or None for a leaf."""
    if isinstance(node, Module):
        cls = type(node)
        static = _static_fields(cls)
        keys = [k for k in cls.__eqx_fields__ if k not in static and hasattr(node, k)]

        def rebuild(ch, node=node, cls=cls, keys=keys):
            new = cls.__new__(cls)
            for k, v in node.__dict__.items():
                object.__setattr__(new, k, v)
            for k, v in zip(keys, ch):
                object.__setattr__(new, k, v)
            return new

        return "module", keys, [getattr(node, k) for k in keys], rebuild
    if _is_namedtuple(node):
        return "namedtuple", list(type(node)._fields), list(node), lambda ch, t=type(node): t(*ch)
    if isinstance(node, tuple):
        return "tuple", list(range(len(node))), list(node), lambda ch: tuple(ch)
    if isinstance(node, list):
        return "list", list(range(len(node))), list(node), lambda ch: list(ch)
    if isinstance(node, dict):
        keys = list(node.keys())
        return "dict", keys, [node[k] for k in keys], lambda ch, keys=keys: dict(zip(keys, ch))
    return None


def tree_map(f, tree, *rest, is_leaf=None):
    if is_leaf is not None and is_leaf(tree):
        return f(tree, *rest)
    if tree is None:
        return None
    info = _children(tree)
    if info is None:
        return f(tree, *rest)
    kind, keys, ch, rebuild = info
    rest_ch = []
    for r in rest:
        ri = _children(r)
        if ri is None:
            raise ValueError(f"tree_map: structure mismatch at {type(tree).__name__} vs {type(r).__name__}")
        if ri[1] != keys:
            raise ValueError(f"tree_map: structure mismatch ({ri[1]} vs {keys})")
        rest_ch.append(ri[2])
    out = [tree_map(f, c, *[rc[i] for rc in rest_ch], is_leaf=is_leaf) for i, c in enumerate(ch)]
    return rebuild(out)


def tree_leaves(tree, is_leaf=None):
    out = []
    tree_map(lambda x: out.append(x), tree, is_leaf=is_leaf)
    return out


class TreeDef:
    """Structure of a pytree: the tree itself with every leaf replaced by a counter slot."""

    def __init__(self, skeleton, n):
        self.skeleton, self.num_leaves = skeleton, n

    def unflatten(self, leaves):
        leaves = list(leaves)
        return tree_map(lambda slot: leaves[slot.i], self.skeleton)

    def __eq__(self, other):
        return isinstance(other, TreeDef) and self.num_leaves == other.num_leaves and tree_equal(
            tree_map(lambda s: 0, self.skeleton), tree_map(lambda s: 0, other.skeleton))


class _Slot:
    def __init__(self, i):
        self.i = i


def tree_flatten(tree, is_leaf=None):
    leaves = []

    def rec(x):
        leaves.append(x)
        return _Slot(len(leaves) - 1)

    skeleton = tree_map(rec, tree, is_leaf=is_leaf)
    return leaves, TreeDef(skeleton, len(leaves))


def tree_structure(tree, is_leaf=None):
    return tree_flatten(tree, is_leaf)[1]


def tree_unflatten(treedef, leaves):
    return treedef.unflatten(leaves)


def _leaf_equal(a, b):
    if isinstance(a, torch.Tensor) or isinstance(b, torch.Tensor):
        return isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor) and a.shape == b.shape and \
            a.dtype == b.dtype and bool(torch.equal(a, b))
    try:
        return bool(a == b)
    except Exception:
        return a is b


def tree_equal(*trees) -> bool:
    first = trees[0]
    for other in trees[1:]:
        if (first is None) != (other is None):
            return False
        if first is None:
            continue
        ia, ib = _children(first), _children(other)
        if (ia is None) != (ib is None):
            return False
        if ia is None:
            if not _leaf_equal(first, other):
                return False
            continue
        if type(first) is not type(other) or ia[1] != ib[1]:
            return False
        if not all(tree_equal(x, y) for x, y in zip(ia[2], ib[2])):
            return False
    return True


def _spec_fn(filter_spec):
    if callable(filter_spec):
        return filter_spec
    if isinstance(filter_spec, bool):
        return lambda x: filter_spec
    raise NotImplementedError("only callable / bool filter specs are supported by this stand-in")


def partition(pytree, filter_spec, replace=None, is_leaf=None):
    fn = _spec_fn(filter_spec)
    yes = tree_map(lambda x: x if fn(x) else replace, pytree, is_leaf=is_leaf)
    no = tree_map(lambda x: replace if fn(x) else x, pytree, is_leaf=is_leaf)
    return yes, no


def filter(pytree, filter_spec, inverse=False, replace=None, is_leaf=None):  # noqa: A001
    fn = _spec_fn(filter_spec)
    return tree_map(lambda x: x if bool(fn(x)) != inverse else replace, pytree, is_leaf=is_leaf)


def combine(*pytrees, is_leaf=None):
    def pick(*xs):
        for x in xs:
            if x is not None:
                return x
        return None

    none_leaf = (lambda x: x is None) if is_leaf is None else (lambda x: x is None or is_leaf(x))
    return tree_map(pick, pytrees[0], *pytrees[1:], is_leaf=none_leaf)


def tree_at(where, pytree, replace=_MISSING, replace_fn=_MISSING, is_leaf=None):
    """Out-of-place update of the node(s) `where` selects (equinox's sentinel approach: every leaf is wrapped in a
    unique object, `where` is applied to the wrapped copy, and the selected nodes are matched by identity)."""

    class _Wrapped:
        def __init__(self, value):
            self.value = value

    none_leaf = (lambda x: x is None) if is_leaf is None else (lambda x: x is None or is_leaf(x))
    wrapped = tree_map(lambda x: _Wrapped(x), pytree, is_leaf=none_leaf)
    sel = where(wrapped)
    single = not isinstance(sel, (tuple, list))
    nodes = [sel] if single else list(sel)
    if replace is not _MISSING:
        reps = [replace] if single else list(replace)
        if len(reps) != len(nodes):
            raise ValueError("tree_at: `where` and `replace` differ in length")
    ids = {id(n): i for i, n in enumerate(nodes)}

    def unwrap(n):
        return tree_map(lambda w: w.value, n, is_leaf=lambda x: isinstance(x, _Wrapped))

    def rec(w):
        if id(w) in ids:
            i = ids[id(w)]
            return reps[i] if replace is not _MISSING else replace_fn(unwrap(w))
        if isinstance(w, _Wrapped):
            return w.value
        if w is None:
            return None
        info = _children(w)
        if info is None:
            return w
        return info[3]([rec(c) for c in info[2]])

    return rec(wrapped)


# --------------------------------------------------------------------------- transforms
def filter_jit(fn=None, **_):
    if fn is None:
        return lambda f: f
    return fn


def filter_eval_shape(fn, *args, **kwargs):
    return fn(*args, **kwargs)


def _stack_leaves(*xs):
    x0 = xs[0]
    if isinstance(x0, torch.Tensor):
        return torch.stack(list(xs))
    if isinstance(x0, (bool, int, float)):
        from jax import numpy as jnp  # the shim's

        return torch.stack([jnp.asarray(x) for x in xs])
    return x0


def _vmap_call(fn, args, kwargs, arg_axes, kw_axes):
    """Loop over axis 0 of the mapped array leaves, stack the outputs."""
    n = None

    def scan_size(tree, ax):
        nonlocal n
        for leaf in tree_leaves(tree):
            a = ax(leaf) if callable(ax) else ax
            if a is not None and isinstance(leaf, torch.Tensor):
                if a != 0:
                    raise NotImplementedError("only axis 0 is supported")
                n = leaf.shape[0] if n is None else n
                if leaf.shape[0] != n:
                    raise ValueError("vmap: inconsistent batch sizes")

    for a, ax in zip(args, arg_axes):
        scan_size(a, ax)
    for k, a in kwargs.items():
        scan_size(a, kw_axes.get(k, 0))
    if n is None:
        raise ValueError("vmap: nothing to map over")

    def take(tree, ax, i):
        def f(leaf):
            a = ax(leaf) if callable(ax) else ax
            return leaf[i] if (a is not None and isinstance(leaf, torch.Tensor)) else leaf

        return tree_map(f, tree)

    outs = [fn(*[take(a, ax, i) for a, ax in zip(args, arg_axes)],
               **{k: take(a, kw_axes.get(k, 0), i) for k, a in kwargs.items()}) for i in range(n)]
    return tree_map(_stack_leaves, outs[0], *outs[1:])


def filter_vmap(fn=None, *, in_axes=if_array(0), out_axes=if_array(0), **_):
    if fn is None:
        return lambda f: filter_vmap(f, in_axes=in_axes, out_axes=out_axes)

    def wrapped(*args, **kwargs):
        axes = list(in_axes) if isinstance(in_axes, (tuple, list)) else [in_axes] * len(args)
        return _vmap_call(fn, args, kwargs, axes, {})

    return wrapped


def filter_value_and_grad(fn=None, *, has_aux=False):
    if fn is None:
        return lambda f: filter_value_and_grad(f, has_aux=has_aux)

    def wrapped(model, *args, **kwargs):
        leaves, treedef = tree_flatten(model)
        new_leaves, variables, slots = list(leaves), [], []
        for i, leaf in enumerate(leaves):
            if is_inexact_array(leaf):
                v = leaf.detach().clone().requires_grad_(True)
                new_leaves[i] = v
                variables.append(v)
                slots.append(i)
        out = fn(treedef.unflatten(new_leaves), *args, **kwargs)
        value, aux = out if has_aux else (out, None)
        grads = torch.autograd.grad(value, variables, allow_unused=True)
        grad_leaves = [None] * len(leaves)
        for i, g, v in zip(slots, grads, variables):
            grad_leaves[i] = torch.zeros_like(v) if g is None else g.detach()
        detach = lambda t: tree_map(lambda x: x.detach() if isinstance(x, torch.Tensor) else x, t)  # noqa: E731
        grad_tree = treedef.unflatten(grad_leaves)
        return ((detach(value), detach(aux)), grad_tree) if has_aux else (detach(value), grad_tree)

    return wrapped


def filter_grad(fn=None, *, has_aux=False):
    vg = filter_value_and_grad(fn, has_aux=has_aux)

    def wrapped(*a, **k):
        (v, g) = vg(*a, **k)
        return (g, v[1]) if has_aux else g

    return wrapped


def apply_updates(model, updates):
    return tree_map(lambda u, p: p if u is None else p + u, updates, model, is_leaf=lambda x: x is None)


def error_if(x, pred, msg):
    p = pred if isinstance(pred, torch.Tensor) else torch.as_tensor(pred)
    if bool(p.any()):
        raise RuntimeError(msg)
    return x


def tree_serialise_leaves(*a, **k):
    raise NotImplementedError


def tree_deserialise_leaves(*a, **k):
    raise NotImplementedError


from . import nn  # noqa: E402,F401
