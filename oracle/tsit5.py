"""One explicit Tsitouras 5(4) step, as diffrax 0.7.2 takes it for the classic-control envs.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows /root/reference/src/lerax/env/classic_control/base_classic_control.py:49-72:
``diffrax.diffeqsolve(ODETerm(dynamics), Tsit5(), t0=t, t1=t+dt, dt0=dt,
ConstantStepSize(), SaveAt(t1=True))`` is exactly ONE step.  diffrax (third-party,
pinned 0.7.2 in /root/reference/uv.lock:367-368, not vendored) forms
    control h = t1 - t0                      (ODETerm.contr; fp32: h = fl(fl(t+dt) - t))
    k_i = f(y_i) * h
    y_i = y0 + sum_j a_ij k_j                (increments pre-multiplied by h)
    y1  = y0 + sum_j b_j k_j                 (7th FSAL stage only feeds the unused error estimate)
with the Tsitouras (2011) tableau below.  The dynamics here are autonomous, so the
stage times c_i are not needed.
"""

from __future__ import annotations

import numpy as np

# Butcher tableau (Tsitouras 2011, "Runge-Kutta pairs of order 5(4) satisfying only the
# first column simplifying assumption").  Double precision; cast to the working dtype.
C = np.array([0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0])
A = [
    np.array([0.161]),
    np.array([-0.008480655492356989, 0.335480655492357]),
    np.array([2.8971530571054935, -6.359448489975075, 4.3622954328695815]),
    np.array([5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525]),
    np.array([5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383]),
]
B = np.array([0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774])


def step(f, y0, h):
    """y0: [k, ...] array (state components on axis 0); f(y)->same shape; h scalar or [...]"""
    dt = y0.dtype.type
    ks = []
    y = y0
    for i in range(6):
        if i > 0:
            acc = dt(A[i - 1][0]) * ks[0]
            for j in range(1, i):
                acc = acc + dt(A[i - 1][j]) * ks[j]
            y = y0 + acc
        ks.append(f(y) * h)
    acc = dt(B[0]) * ks[0]
    for j in range(1, 6):
        acc = acc + dt(B[j]) * ks[j]
    return y0 + acc


def step_size(t, dt):
    """h = t1 - t0 with t1 = t + dt, evaluated in the working precision."""
    return (t + dt) - t
