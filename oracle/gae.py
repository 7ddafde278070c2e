"""Generalised Advantage Estimation.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates /root/reference/src/lerax/advantage/gae.py:36-66 (called from
buffer/rollout.py:64-79 and algorithm/ppo.py:311-315):
    next_v = [v_1 .. v_{T-1}, last_value];  nd = 1 - done
    delta  = r + gamma * next_v * nd - v ;  disc = gamma * lam * nd
    A_t    = delta_t + disc_t * A_{t+1}   (reverse scan, carry 0) ; returns = A + v
Pinned against the reference's known-answer tests tests/advantage/test_gae.py:22-85.
"""

from __future__ import annotations

import numpy as np


def gae(rewards, values, dones, last_value, gamma, lam):
    """Time on axis 0; any trailing (env) axes are vectorised.  Returns (adv, ret)."""
    dt = rewards.dtype.type
    gamma, lam = dt(gamma), dt(lam)
    T = rewards.shape[0]
    last_value = np.asarray(last_value, dtype=rewards.dtype)
    next_values = np.concatenate([values[1:], last_value[None]], axis=0)
    not_done = dt(1.0) - dones.astype(rewards.dtype)
    deltas = rewards + gamma * next_values * not_done - values
    discounts = gamma * lam * not_done
    adv = np.empty_like(rewards)
    carry = np.zeros(rewards.shape[1:], dtype=rewards.dtype)
    for t in range(T - 1, -1, -1):
        carry = deltas[t] + discounts[t] * carry
        adv[t] = carry
    return adv, adv + values


def gae_loop_reference(rewards, values, dones, last_value, gamma, lam):
    """The in-test loop of tests/advantage/test_gae.py:8-19, for 1-D inputs."""
    T = rewards.shape[0]
    adv = np.zeros_like(rewards)
    nxt = rewards.dtype.type(0)
    for i in range(T - 1, -1, -1):
        nv = last_value if i == T - 1 else values[i + 1]
        nd = rewards.dtype.type(1.0) - rewards.dtype.type(dones[i])
        delta = rewards[i] + rewards.dtype.type(gamma) * nv * nd - values[i]
        nxt = delta + rewards.dtype.type(gamma) * rewards.dtype.type(lam) * nd * nxt
        adv[i] = nxt
    return adv, adv + values
