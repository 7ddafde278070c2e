"""Categorical and Normal, as distreqx 0.0.2 computes them behind the Lerax wrappers.

TEST INFRASTRUCTURE (see oracle/__init__.py).

/root/reference/src/lerax/distribution/categorical.py:12-50 and normal.py:12-46 delegate
through base_distribution.py:47-78 to distreqx 0.0.2 (third-party,
/root/reference/uv.lock:385-386, not vendored).  Published algorithm restated:

  Categorical(logits): logits <- log_softmax(logits); log_prob(a) = logits[a];
      entropy = -sum_k exp(logits_k) * logits_k (0 where logits_k = -inf);
      sample = argmax_k(logits_k + Gumbel_k)  [jax.random.categorical]; mode = argmax.
  Normal(loc, scale): log_prob(x) = -0.5*((x-loc)/scale)^2 - (0.5*log(2*pi) + log(scale));
      entropy = 0.5 + 0.5*log(2*pi) + log(scale); sample = loc + scale*eps.

Pinned by the reference's own tests: tests/distribution/test_categorical.py:60-69,82-96,
tests/distribution/test_normal.py:49-65,75-82 (re-expressed in tests/test_oracle_golden.py).
"""

from __future__ import annotations

import numpy as np


def log_softmax(logits):
    m = np.max(logits, axis=-1, keepdims=True)
    s = logits - m
    return s - np.log(np.sum(np.exp(s), axis=-1, keepdims=True))


def categorical_log_prob(logits, action):
    lp = log_softmax(logits)
    return np.take_along_axis(lp, action[..., None].astype(np.int64), axis=-1)[..., 0]


def categorical_entropy(logits):
    lp = log_softmax(logits)
    p = np.exp(lp)
    with np.errstate(invalid="ignore"):
        term = np.where(np.isneginf(lp), 0, p * lp)
    return -np.sum(term, axis=-1)


def categorical_sample(logits, gumbel):
    """argmax(logits + g); ties -> lowest index (jnp.argmax)."""
    return np.argmax(logits + gumbel, axis=-1).astype(np.int32)


def categorical_mode(logits):
    return np.argmax(logits, axis=-1).astype(np.int32)


def normal_log_prob(loc, scale, x):
    dt = np.asarray(loc).dtype.type
    z = (x - loc) / scale
    return dt(-0.5) * (z * z) - (dt(0.5 * np.log(2 * np.pi)) + np.log(scale))


def normal_entropy(scale):
    dt = np.asarray(scale).dtype.type
    return dt(0.5) + dt(0.5 * np.log(2 * np.pi)) + np.log(scale)


def normal_sample(loc, scale, eps):
    return loc + scale * eps
