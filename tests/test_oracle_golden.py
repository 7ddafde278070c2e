"""Pin the oracle against the reference's own known-answer tests (tests/golden/*.json,
transcribed by tests/golden/make_golden.py from /root/reference/tests) and against
published constants (Philox KAT, Tsit5 order conditions)."""

import json
import os

import numpy as np
import pytest

from oracle import distributions as D
from oracle import envs as E
from oracle import gae as G
from oracle import philox, tsit5

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


# ------------------------------------------------------------------ GAE (pinned)
@pytest.mark.parametrize("case", load("gae.json"), ids=lambda c: c["name"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_gae_golden(case, dtype):
    r = np.array(case["rewards"], dtype)
    v = np.array(case["values"], dtype)
    d = np.array(case["dones"], bool)
    adv, ret = G.gae(r, v, d, dtype(case["last_value"]), case["gamma"], case["lam"])
    tol = 1e-6 if dtype == np.float32 else 1e-12
    np.testing.assert_allclose(adv, case["advantages"], rtol=tol, atol=tol)
    np.testing.assert_allclose(ret, case["returns"], rtol=tol, atol=tol)
    adv2, ret2 = G.gae_loop_reference(r, v, d, dtype(case["last_value"]), case["gamma"], case["lam"])
    np.testing.assert_allclose(adv, adv2, rtol=tol, atol=tol)
    np.testing.assert_allclose(ret, adv + v, rtol=0, atol=0)  # test_buffer.py:82 returns = adv + values


def test_gae_vectorised_matches_loop():
    rng = np.random.default_rng(4)
    T, N = 37, 11
    r = rng.standard_normal((T, N)).astype(np.float32)
    v = rng.standard_normal((T, N)).astype(np.float32)
    d = rng.random((T, N)) < 0.1
    last = rng.standard_normal(N).astype(np.float32)
    adv, ret = G.gae(r, v, d, last, 0.99, 0.95)
    for n in range(N):
        a1, r1 = G.gae_loop_reference(r[:, n], v[:, n], d[:, n], last[n], 0.99, 0.95)
        np.testing.assert_allclose(adv[:, n], a1, rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(ret[:, n], r1, rtol=1e-6, atol=1e-6)


# -------------------------------------------------------- distributions (pinned)
def test_categorical_golden():
    for c in load("distributions.json")["categorical"]:
        with np.errstate(divide="ignore"):
            logits = np.log(np.array(c["probs"], np.float32))  # distreqx: probs -> log(probs)
        if "log_probs" in c:
            lp = D.categorical_log_prob(np.tile(logits, (len(c["actions"]), 1)), np.array(c["actions"]))
            np.testing.assert_allclose(lp, c["log_probs"], rtol=1e-6)
        if "entropy" in c:
            np.testing.assert_allclose(D.categorical_entropy(logits), c["entropy"], rtol=1e-6, atol=1e-6)
        if "mode" in c:
            assert D.categorical_mode(logits) == c["mode"]


def test_categorical_sample_and_log_prob_consistent():
    # test_categorical.py:49-58
    logits = np.array([[0.0, 1.0, 2.0]], np.float32)
    g = philox.gumbels(0, np.arange(1), np.zeros(1, np.uint32), 3)
    a = D.categorical_sample(D.log_softmax(logits), g)
    assert 0 <= a[0] < 3
    np.testing.assert_allclose(D.categorical_log_prob(logits, a), D.log_softmax(logits)[0, a[0]])


def test_normal_golden():
    for c in load("distributions.json")["normal"]:
        loc, scale = np.float32(c["loc"]), np.float32(c["scale"])
        if "log_prob" in c:
            np.testing.assert_allclose(D.normal_log_prob(loc, scale, np.float32(c["x"])), c["log_prob"], rtol=1e-6)
        if "entropy" in c:
            np.testing.assert_allclose(D.normal_entropy(scale), c["entropy"], rtol=1e-6)


# ------------------------------------------------------------------ Philox (KAT)
def test_philox_kat():
    for c in load("philox.json"):
        out = philox.philox4x32_10(*[np.uint32(x) for x in c["counter"]], *c["key"])
        assert [int(o) for o in out] == c["out"]


def test_noise_ranges():
    env = np.arange(4096, dtype=np.uint32)
    u = philox.uniforms(123, env, np.uint32(7), philox.STREAM_RESET, 4)
    assert u.dtype == np.float32 and (u > 0).all() and (u < 1).all()
    assert abs(u.mean() - 0.5) < 0.02
    z = philox.normals(123, env, np.uint32(7))
    assert abs(z.mean()) < 0.06 and abs(z.std() - 1) < 0.06
    g = philox.gumbels(123, env, np.uint32(7), 3)
    assert abs(g.mean() - 0.5772) < 0.06


# -------------------------------------------------- Tsit5 (published conditions)
def test_tsit5_order_conditions():
    c = tsit5.C[:6]
    b = tsit5.B
    A = np.zeros((6, 6))
    for i, row in enumerate(tsit5.A):
        A[i + 1, : len(row)] = row
    np.testing.assert_allclose(A.sum(1), c, atol=1e-14)  # row sums
    assert abs(b.sum() - 1) < 1e-14
    for k in range(1, 5):
        assert abs(b @ c**k - 1 / (k + 1)) < 1e-14
    assert abs(b @ (A @ c) - 1 / 6) < 1e-14
    assert abs(b @ (c * (A @ c)) - 1 / 8) < 1e-14
    assert abs(b @ (A @ c**2) - 1 / 12) < 1e-14
    assert abs(b @ (A @ (A @ c)) - 1 / 24) < 1e-14
    assert abs(b @ (A @ (A @ (A @ c))) - 1 / 120) < 1e-14


def test_tsit5_reproduces_taylor_series_to_fifth_order():
    """On y' = y one step must equal the Taylor series of e^h through h^5 (a 5th-order
    method); Tsit5's h^6 error constant is small, so |y1 - e^h| << h^6/720."""
    for h in (0.4, 0.2, 0.1):
        y1 = tsit5.step(lambda y: y, np.array([[1.0]]), h)[0, 0]
        assert abs(y1 - np.exp(h)) < 0.2 * h**6 / 720, h


def test_tsit5_pendulum_step_vs_scipy_dop853():
    """One env-sized step (dt = 0.05) against a tight DOP853 solution of the same ODE."""
    from scipy.integrate import solve_ivp

    spec = E.pendulum()
    y0 = np.array([[0.7], [-0.4]])
    act = np.array([1.3])
    errs = []
    for h in (0.1, 0.05):
        y1 = tsit5.step(lambda yy: E.dynamics(spec, yy, act), y0, h)
        sol = solve_ivp(lambda t, y: E.dynamics(spec, y[:, None], act)[:, 0], (0, h), y0[:, 0],
                        method="DOP853", rtol=1e-13, atol=1e-15)
        errs.append(np.abs(y1[:, 0] - sol.y[:, -1]).max())
    assert errs[0] < 1e-7 and errs[1] < 5e-9 and errs[1] < errs[0] / 16, errs


# ------------------------------------------------------------- env closed forms
def test_cartpole_euler_limit_matches_gymnasium_formula():
    """CartPole docstring (cartpole.py:38-40): identical dynamics to Gymnasium; check the
    derivative against the Gymnasium closed form at a hand-picked state."""
    spec = E.cartpole()
    y = np.array([[0.1], [0.5], [0.05], [-0.3]])
    dy = E.dynamics(spec, y, np.array([1]))
    g, mc, mp, l, F = 9.8, 1.0, 0.1, 0.5, 10.0
    th, thd = 0.05, -0.3
    temp = (F + mp * l * thd**2 * np.sin(th)) / (mc + mp)
    thacc = (g * np.sin(th) - np.cos(th) * temp) / (l * (4 / 3 - mp * np.cos(th) ** 2 / (mc + mp)))
    xacc = temp - mp * l * thacc * np.cos(th) / (mc + mp)
    np.testing.assert_allclose(dy[:, 0], [0.5, xacc, -0.3, thacc], rtol=1e-7)  # params are fp32-rounded like jnp.array(0.1)


def test_cartpole_terminal_bounds_inclusive():
    spec = E.cartpole()
    xt = np.float32(2.4)
    tt = np.float32(12 * 2 * np.pi / 360)
    y = np.zeros((4, 4), np.float32)
    y[0] = [xt, np.nextafter(xt, np.float32(3)), 0, 0]
    y[2] = [0, 0, tt, np.nextafter(tt, np.float32(1))]
    assert E.terminal(spec, y).tolist() == [False, True, False, True]


def test_pendulum_clip_and_reward():
    spec = E.pendulum()
    y = np.array([[4.0, -4.0, 0.5], [9.0, -9.0, 1.0]], np.float32)
    c = E.clip_state(spec, y)
    assert (c[0] >= -np.pi).all() and (c[0] < np.pi).all()
    np.testing.assert_allclose(c[0], [4.0 - 2 * np.pi, -4.0 + 2 * np.pi, 0.5], rtol=1e-6)
    assert c[1].tolist() == [8.0, -8.0, 1.0]
    r = E.reward(spec, y, np.array([3.0, -3.0, 0.5], np.float32), c)
    np.testing.assert_allclose(r, -(c[0] ** 2 + 0.1 * c[1] ** 2 + 0.001 * np.array([4, 4, 0.25])), rtol=1e-6)


def test_acrobot_reward_terminal():
    spec = E.acrobot()
    y = np.array([[np.pi, 0.0], [0.0, 0.0], [0, 0], [0, 0]], np.float32)
    assert E.terminal(spec, y).tolist() == [True, False]
    assert E.reward(spec, y, np.array([0, 0]), y).tolist() == [0.0, -1.0]


def test_initial_ranges():
    u = np.random.default_rng(1).random((4, 100)).astype(np.float32)
    y = E.initial_from_uniform(E.cartpole(), u)
    assert (np.abs(y) <= 0.05).all()
    y = E.initial_from_uniform(E.pendulum(), u[:2])
    assert (np.abs(y[0]) <= np.pi + 1e-6).all() and (np.abs(y[1]) <= 1).all()
    y = E.initial_from_uniform(E.acrobot(), u)
    assert (np.abs(y) <= 0.1 + 1e-7).all()


# ------------------------------------------------------------------ MountainCar / ContinuousMountainCar
def test_mountain_car_semantics():
    """mountain_car.py:145-186: initial range, constant -1 reward, left-wall stop, goal test."""
    spec = E.mountain_car()
    u = np.array([[0.0, 0.5, 0.999999], [0.3, 0.3, 0.3]], np.float32)
    y = E.initial_from_uniform(spec, u)
    assert np.allclose(y[0], [-0.6, -0.5, -0.4], atol=1e-6) and np.all(y[1] == 0)
    act = np.array([0, 1, 2])
    y1, t1 = E.transition(spec, y, np.zeros(3, np.float32), act)
    assert np.all(t1 == 1.0) and np.all(E.reward(spec, y, act, y1) == -1.0)
    assert y1[1, 0] < y1[1, 1] < y1[1, 2]  # push left < none < push right
    # at the left wall a negative velocity is zeroed, a positive one survives (the car stops dead)
    wall = E.clip_state(spec, np.array([[-1.3, -1.3, -1.0], [-0.05, 0.03, -0.2]], np.float32))
    assert np.allclose(wall, [[-1.2, -1.2, -1.0], [0.0, 0.03, -0.07]])
    term = E.terminal(spec, np.array([[0.5, 0.5, 0.49], [0.0, -0.01, 0.05]], np.float32))
    assert term.tolist() == [True, False, False]


def test_continuous_mountain_car_semantics():
    """continuous_mountain_car.py:132-174: action clipped to the Box, reward -0.1 a^2 (+100 when the PRE-transition
    state is at the goal), no wall stop."""
    spec = E.continuous_mountain_car()
    assert E.action_bounds(spec) == (np.float32(-1.0), np.float32(1.0))
    y = np.array([[-0.5, -0.5, 0.55], [0.0, 0.0, 0.01]], np.float32)
    act = np.array([3.0, -0.5, 0.2], np.float32)
    y1, _ = E.transition(spec, y, np.zeros(3, np.float32), act)
    r = E.reward(spec, y, act, y1)
    assert np.allclose(r, [-0.1, -0.025, 100.0 - 0.1 * 0.04], rtol=1e-6)
    y_sat, _ = E.transition(spec, y, np.zeros(3, np.float32), np.array([1.0, -0.5, 0.2], np.float32))
    assert y1[1, 0] == y_sat[1, 0]  # 3.0 acts like 1.0
    wall = E.clip_state(spec, np.array([[-1.3], [-0.05]], np.float32))
    assert np.allclose(wall, [[-1.2], [-0.05]])
