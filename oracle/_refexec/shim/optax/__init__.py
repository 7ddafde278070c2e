"""Stand-in for optax 0.2.6 (uv.lock:1783-1784) -- TEST INFRASTRUCTURE, see oracle/_refexec/__init__.py.
chain(clip_by_global_norm(c), inject_hyperparams(adam)(lr)) as algorithm/ppo.py:192-194 builds it:
  clip_by_global_norm: g_norm = sqrt(sum g^2); g if g_norm < c else (g / g_norm) * c
  scale_by_adam: mu = b1 mu + (1-b1) g ; nu = b2 nu + (1-b2) g^2 ; count += 1 ;
                 mu_hat = mu / (1 - b1^count) ; nu_hat = nu / (1 - b2^count) ; u = mu_hat / (sqrt(nu_hat + eps_root) + eps)
  scale_by_learning_rate: u * -lr      (lr may be a schedule of the adam count, via inject_hyperparams)."""

from __future__ import annotations

from typing import Any, Callable, NamedTuple, Union

import equinox as eqx
import torch

OptState = Any
ScalarOrSchedule = Union[float, Callable]
Params = Updates = Any


class GradientTransformation(NamedTuple):
    init: Callable
    update: Callable


class EmptyState(NamedTuple):
    pass


class ScaleByAdamState(NamedTuple):
    count: torch.Tensor
    mu: Any
    nu: Any


class InjectHyperparamsState(NamedTuple):
    count: torch.Tensor
    hyperparams: dict
    inner_state: Any


def _map(f, *trees):
    return eqx.tree_map(f, *trees)


def global_norm(updates):
    leaves = [x for x in eqx.tree_leaves(updates) if isinstance(x, torch.Tensor)]
    return torch.sqrt(sum((x * x).sum() for x in leaves))


def clip_by_global_norm(max_norm):
    def init(params):
        return EmptyState()

    def update(updates, state, params=None):
        g_norm = global_norm(updates)
        trigger = g_norm < max_norm
        return _map(lambda t: torch.where(trigger, t, (t / g_norm) * max_norm), updates), state

    return GradientTransformation(init, update)


def _safe_int32_increment(count):
    return count + 1


def scale_by_adam(b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0):
    def init(params):
        z = _map(torch.zeros_like, params)
        return ScaleByAdamState(torch.zeros((), dtype=torch.int32), z, _map(torch.zeros_like, params))

    def update(updates, state, params=None):
        mu = _map(lambda g, m: (1 - b1) * g + b1 * m, updates, state.mu)
        nu = _map(lambda g, v: (1 - b2) * (g * g) + b2 * v, updates, state.nu)
        count = _safe_int32_increment(state.count)
        c = count.to(torch.float32)
        bc1, bc2 = 1 - torch.tensor(b1, dtype=torch.float32) ** c, 1 - torch.tensor(b2, dtype=torch.float32) ** c
        out = _map(lambda m, v: (m / bc1) / (torch.sqrt(v / bc2 + eps_root) + eps), mu, nu)
        return out, ScaleByAdamState(count, mu, nu)

    return GradientTransformation(init, update)


def scale_by_learning_rate(learning_rate):
    def init(params):
        return EmptyState()

    def update(updates, state, params=None):
        return _map(lambda u: u * (-learning_rate), updates), state

    return GradientTransformation(init, update)


def chain(*transforms):
    def init(params):
        return tuple(t.init(params) for t in transforms)

    def update(updates, state, params=None):
        new = []
        for t, s in zip(transforms, state):
            updates, s2 = t.update(updates, s, params)
            new.append(s2)
        return updates, tuple(new)

    return GradientTransformation(init, update)


def adam(learning_rate, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0):
    return chain(scale_by_adam(b1, b2, eps, eps_root), scale_by_learning_rate(learning_rate))


def inject_hyperparams(inner_factory):
    """Numeric hyper-parameters live in the state; callables are schedules evaluated at the state's own count."""

    def wrapped(learning_rate, **kw):
        def current(count):
            lr = learning_rate(count) if callable(learning_rate) else learning_rate
            return torch.as_tensor(lr, dtype=torch.float32)

        def init(params):
            count = torch.zeros((), dtype=torch.int32)
            lr = current(count)
            return InjectHyperparamsState(count, {"learning_rate": lr}, inner_factory(lr, **kw).init(params))

        def update(updates, state, params=None):
            lr = current(state.count)
            updates, inner = inner_factory(lr, **kw).update(updates, state.inner_state, params)
            return updates, InjectHyperparamsState(state.count + 1, {"learning_rate": lr}, inner)

        return GradientTransformation(init, update)

    return wrapped


def apply_updates(params, updates):
    return eqx.apply_updates(params, updates)
