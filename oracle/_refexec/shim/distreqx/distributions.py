"""distreqx.distributions.{Categorical, Normal, MultivariateNormalDiag}.

Categorical (distreqx/distributions/categorical.py): logits are normalised at construction
(logits - logsumexp(logits)); log_prob(a) = normalised logits[a]; probs = softmax(logits);
entropy = -sum p * log p with 0 log 0 = 0; sample = jax.random.categorical (Gumbel argmax); mode = argmax.
Normal (normal.py): log_prob = -0.5 z^2 - log(scale) - 0.5 log(2 pi), z = (x - loc) / scale;
entropy = 0.5 + 0.5 log(2 pi) + log(scale); sample = loc + scale * N(0,1).
sample_and_log_prob = sample, then log_prob(sample) (distribution base class)."""

from __future__ import annotations

import math

import equinox as eqx
import torch
from jax import numpy as jnp
from jax import random as jr


class AbstractDistribution(eqx.Module):
    def sample_and_log_prob(self, key):
        s = self.sample(key)
        return s, self.log_prob(s)

    def prob(self, value):
        return torch.exp(self.log_prob(value))


class AbstractTransformed(AbstractDistribution):
    pass


class Categorical(AbstractDistribution):
    _logits: torch.Tensor
    _probs: torch.Tensor

    def __init__(self, logits=None, probs=None):
        if (logits is None) == (probs is None):
            raise ValueError("exactly one of logits / probs")
        if logits is not None:
            lg = jnp.asarray(logits)
            self._logits = lg - torch.logsumexp(lg, dim=-1, keepdim=True)
            self._probs = None
        else:
            p = jnp.asarray(probs)
            self._probs = p / p.sum(dim=-1, keepdim=True)
            self._logits = None

    @property
    def logits(self):
        return self._logits if self._logits is not None else torch.log(self._probs)

    @property
    def probs(self):
        return self._probs if self._probs is not None else torch.softmax(self._logits, dim=-1)

    def log_prob(self, value):
        v = jnp.asarray(value).to(torch.int64)
        return torch.gather(self.logits, -1, v.reshape(self.logits.shape[:-1] + (1,))).squeeze(-1)

    def entropy(self):
        lg, p = self.logits, self.probs
        term = torch.where(p == 0, torch.zeros_like(p), p * lg)
        return -term.sum(dim=-1)

    def sample(self, key):
        return jr.categorical(key, self.logits)

    def mode(self):
        return torch.argmax(self.logits, dim=-1).to(torch.int32)

    def mean(self):
        raise NotImplementedError


class Normal(AbstractDistribution):
    loc: torch.Tensor
    scale: torch.Tensor

    def __init__(self, loc, scale):
        self.loc, self.scale = jnp.asarray(loc), jnp.asarray(scale)

    def log_prob(self, value):
        z = (jnp.asarray(value) - self.loc) / self.scale
        return -0.5 * (z * z) - (0.5 * math.log(2.0 * math.pi) + torch.log(self.scale))

    def entropy(self):
        return 0.5 + 0.5 * math.log(2.0 * math.pi) + torch.log(self.scale)

    def sample(self, key):
        return self.loc + self.scale * jr.normal(key, self.loc.shape)

    def mode(self):
        return self.loc

    def mean(self):
        return self.loc


class MultivariateNormalDiag(AbstractDistribution):
    loc: torch.Tensor
    scale_diag: torch.Tensor

    def __init__(self, loc=None, scale_diag=None):
        self.loc, self.scale_diag = jnp.asarray(loc), jnp.asarray(scale_diag)

    def log_prob(self, value):
        z = (jnp.asarray(value) - self.loc) / self.scale_diag
        return (-0.5 * (z * z) - (0.5 * math.log(2.0 * math.pi) + torch.log(self.scale_diag))).sum(-1)

    def entropy(self):
        return (0.5 + 0.5 * math.log(2.0 * math.pi) + torch.log(self.scale_diag)).sum(-1)

    def sample(self, key):
        return self.loc + self.scale_diag * jr.normal(key, self.loc.shape)

    def mode(self):
        return self.loc

    def mean(self):
        return self.loc


class _Unsupported(AbstractDistribution):
    def __init__(self, *a, **k):
        raise NotImplementedError("not on the on-policy classic-control path")


Bernoulli = Transformed = Independent = _Unsupported
