"""The oracle's BootstrappedReturn / NStepReturn against the reference's own known-answer tests
(/root/reference/tests/advantage/test_bootstrapped.py, test_n_step.py), restated with NumPy inputs."""
import numpy as np
import pytest

from oracle import gae as G

f32 = np.float32


def test_discounted_returns_bootstraps_and_stops_at_episode_boundaries():  # test_bootstrapped.py:8-17
    ret = G.discounted_returns(np.array([1, 2, 3, 4], f32), np.array([False, True, False, False]), f32(5.0), 1.0)
    assert np.allclose(ret, [3.0, 2.0, 12.0, 9.0])


def test_bootstrapped_return_matches_gae_lambda_one():  # test_bootstrapped.py:20-30
    r, v = np.array([1.0, -0.5, 2.0, 0.25], f32), np.array([0.3, 0.1, -0.2, 0.4], f32)
    d, lv = np.array([False, True, False, False]), f32(0.75)
    want = G.gae(r, v, d, lv, 0.9, 1.0)
    got = G.bootstrapped_return(r, v, d, lv, 0.9)
    assert np.allclose(got[0], want[0]) and np.allclose(got[1], want[1])


def test_n_step_one_matches_gae_lambda_zero():  # test_n_step.py:7-17
    r, v = np.array([1, 2, 3], f32), np.array([0.5, 0.25, 1.0], f32)
    d, lv = np.array([False, False, True]), f32(4.0)
    want = G.gae(r, v, d, lv, 0.9, 0.0)
    got = G.n_step_return(r, v, d, lv, 0.9, 1)
    assert np.allclose(got[0], want[0]) and np.allclose(got[1], want[1])


def test_large_n_step_matches_bootstrapped_return():  # test_n_step.py:20-30
    r, v = np.array([1.0, -0.5, 2.0, 0.25], f32), np.array([0.3, 0.1, -0.2, 0.4], f32)
    d, lv = np.array([False, True, False, False]), f32(0.75)
    want = G.bootstrapped_return(r, v, d, lv, 0.9)
    got = G.n_step_return(r, v, d, lv, 0.9, 10)
    assert np.allclose(got[0], want[0]) and np.allclose(got[1], want[1])


def test_n_step_stops_at_episode_boundary():  # test_n_step.py:33-40
    _, ret = G.n_step_return(np.ones(4, f32), np.zeros(4, f32), np.array([False, True, False, False]), f32(5.0), 1.0, 3)
    assert np.allclose(ret, [2.0, 1.0, 7.0, 6.0])


def test_n_step_requires_positive_horizon():  # test_n_step.py:43-45
    with pytest.raises(ValueError, match="n must be at least 1"):
        G.n_step_return(np.ones(2, f32), np.zeros(2, f32), np.zeros(2, bool), f32(0), 0.99, 0)


def test_batched_axes_are_vectorised():
    rng = np.random.default_rng(0)
    r, v = rng.normal(size=(9, 5)).astype(f32), rng.normal(size=(9, 5)).astype(f32)
    d, lv = rng.random((9, 5)) < 0.3, rng.normal(size=5).astype(f32)
    for n in (1, 3, 20):
        a, ret = G.n_step_return(r, v, d, lv, 0.97, n)
        for e in range(5):
            a1, r1 = G.n_step_return(r[:, e], v[:, e], d[:, e], lv[e], 0.97, n)
            assert np.array_equal(a[:, e], a1) and np.array_equal(ret[:, e], r1)
    a, ret = G.bootstrapped_return(r, v, d, lv, 0.97)
    for e in range(5):
        a1, r1 = G.bootstrapped_return(r[:, e], v[:, e], d[:, e], lv[e], 0.97)
        assert np.array_equal(a[:, e], a1) and np.array_equal(ret[:, e], r1)
