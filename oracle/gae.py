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


# ----------------------------------------------------------------------------- other estimators
def discounted_returns(rewards, dones, final_value, gamma):
    """advantage/bootstrapped.py:12-49: reverse scan  G_t = r_t + gamma * (1 - done_t) * G_{t+1},  G_T = final_value.
    Time on axis 0; trailing (env) axes are vectorised.  Pinned against tests/advantage/test_bootstrapped.py:8-17."""
    dt = rewards.dtype.type
    gamma = dt(gamma)
    not_done = dt(1.0) - dones.astype(rewards.dtype)
    carry = np.broadcast_to(np.asarray(final_value, rewards.dtype), rewards.shape[1:]).copy()
    out = np.empty_like(rewards)
    for t in range(rewards.shape[0] - 1, -1, -1):
        carry = rewards[t] + gamma * not_done[t] * carry
        out[t] = carry
    return out


def bootstrapped_return(rewards, values, dones, last_value, gamma):
    """BootstrappedReturn.__call__ (advantage/bootstrapped.py:66-79): (returns - values, returns)."""
    ret = discounted_returns(rewards, dones, last_value, gamma)
    return ret - values, ret


def n_step_return(rewards, values, dones, last_value, gamma, n):
    """NStepReturn.__call__ (advantage/n_step.py:30-70): n rewards, then bootstrap from the value n steps ahead
    (last_value past the segment); an episode boundary zeroes the discount.  The reference evaluates
    carry <- r[t+k] + disc[t+k] * carry for k = n-1 .. 0 over arrays padded with r = 0, disc = 1.
    Pinned against tests/advantage/test_n_step.py:7-42."""
    if n < 1:
        raise ValueError(f"n must be at least 1, got {n}")
    dt = rewards.dtype.type
    T = rewards.shape[0]
    tail = rewards.shape[1:]
    discounts = dt(gamma) * (dt(1.0) - dones.astype(rewards.dtype))
    pr = np.concatenate([rewards, np.zeros((n,) + tail, rewards.dtype)])
    pd = np.concatenate([discounts, np.ones((n,) + tail, rewards.dtype)])
    lv = np.broadcast_to(np.asarray(last_value, values.dtype), (n,) + tail)
    carry = np.concatenate([values, lv])[n:n + T].copy()
    for k in range(n - 1, -1, -1):
        carry = pr[k:k + T] + pd[k:k + T] * carry
    return carry - values, carry
