"""Stand-in for diffrax 0.7.2 (uv.lock:367-368) -- TEST INFRASTRUCTURE, see oracle/_refexec/__init__.py.

The only use (env/classic_control/base_classic_control.py:52-67): diffeqsolve(ODETerm(dynamics), Tsit5(),
t0=t, t1=t+dt, dt0=dt, ConstantStepSize(), SaveAt(t1=True)) = exactly ONE explicit Runge-Kutta step.
diffrax's AbstractERK forms the control h = t1 - t0, increments k_i = f(t0 + c_i h, y_i) * h,
stage values y_i = y0 + sum_j a_ij k_j, and (Tsit5 is FSAL with the solution equal to the last stage)
y1 = y0 + sum_j b_j k_j.  Tableau: Tsitouras (2011), restated here independently of oracle/tsit5.py (its order
conditions are checked in tests/test_oracle_golden.py)."""

from __future__ import annotations

import torch

_C = (0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0)
_A = (
    (),
    (0.161,),
    (-0.008480655492356989, 0.335480655492357),
    (2.8971530571054935, -6.359448489975075, 4.3622954328695815),
    (5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525),
    (5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383),
)
_B = (0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774)


class AbstractSolver:
    pass


class AbstractStepSizeController:
    pass


class AbstractAdaptiveStepSizeController(AbstractStepSizeController):
    pass


class ConstantStepSize(AbstractStepSizeController):
    pass


class Tsit5(AbstractSolver):
    def step(self, term, t0, t1, y0, args):
        h = t1 - t0
        ks = []
        for i in range(6):
            y = y0
            if i > 0:
                acc = _A[i][0] * ks[0]
                for j in range(1, i):
                    acc = acc + _A[i][j] * ks[j]
                y = y0 + acc
            ks.append(term.vf(t0 + _C[i] * h, y, args) * h)
        acc = _B[0] * ks[0]
        for j in range(1, 6):
            acc = acc + _B[j] * ks[j]
        return y0 + acc


class Euler(AbstractSolver):
    def step(self, term, t0, t1, y0, args):
        return y0 + term.vf(t0, y0, args) * (t1 - t0)


class ODETerm:
    def __init__(self, vector_field):
        self.vf = vector_field


class SaveAt:
    def __init__(self, t0=False, t1=False, ts=None, steps=False):
        if not t1 or t0 or ts is not None or steps:
            raise NotImplementedError("only SaveAt(t1=True)")


class Solution:
    def __init__(self, ts, ys):
        self.ts, self.ys = ts, ys


def diffeqsolve(terms, solver, t0, t1, dt0, y0, args=None, *, saveat=None, stepsize_controller=None, **_):
    if not isinstance(stepsize_controller, ConstantStepSize):
        raise NotImplementedError("only ConstantStepSize")
    t0, t1, dt0 = (torch.as_tensor(v, dtype=torch.float32) for v in (t0, t1, dt0))
    # one step iff t0 + dt0 reaches t1 (the envs pass dt0 = dt and t1 = t + dt)
    if not bool(t0 + dt0 >= t1):
        raise NotImplementedError("stand-in supports exactly one step")
    y1 = solver.step(terms, t0, t1, y0, args)
    return Solution(t1[None], y1[None])
