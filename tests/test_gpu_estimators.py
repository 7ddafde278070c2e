"""lrx_returns (BootstrappedReturn / NStepReturn) against the oracle and the reference's known answers."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
f32 = np.float32


def _t(x):
    import torch

    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def test_reference_known_answers():
    """tests/advantage/test_bootstrapped.py:8-30 and test_n_step.py:7-45 of the reference, through the GPU classes."""
    from lerax_b200.advantage import GAE, BootstrappedReturn, NStepReturn, discounted_returns

    ret = discounted_returns(_t(np.array([1, 2, 3, 4], f32)), _t(np.array([False, True, False, False])), 5.0, 1.0)
    assert np.allclose(ret.cpu().numpy(), [3.0, 2.0, 12.0, 9.0])
    r, v = _t(np.array([1.0, -0.5, 2.0, 0.25], f32)), _t(np.array([0.3, 0.1, -0.2, 0.4], f32))
    d = _t(np.array([False, True, False, False]))
    want = GAE(gamma=0.9, lam=1.0)(r, v, d, 0.75)
    got = BootstrappedReturn(gamma=0.9)(r, v, d, 0.75)
    big = NStepReturn(n=10, gamma=0.9)(r, v, d, 0.75)
    for i in range(2):
        assert np.allclose(got[i].cpu().numpy(), want[i].cpu().numpy())
        assert np.allclose(big[i].cpu().numpy(), got[i].cpu().numpy())
    r3, v3, d3 = _t(np.array([1, 2, 3], f32)), _t(np.array([0.5, 0.25, 1.0], f32)), _t(np.array([False, False, True]))
    want = GAE(gamma=0.9, lam=0.0)(r3, v3, d3, 4.0)
    got = NStepReturn(n=1, gamma=0.9)(r3, v3, d3, 4.0)
    for i in range(2):
        assert np.allclose(got[i].cpu().numpy(), want[i].cpu().numpy())
    _, ret = NStepReturn(n=3, gamma=1.0)(_t(np.ones(4, f32)), _t(np.zeros(4, f32)), d, 5.0)
    assert np.allclose(ret.cpu().numpy(), [2.0, 1.0, 7.0, 6.0])
    with pytest.raises(ValueError, match="n must be at least 1"):
        NStepReturn(n=0)


@pytest.mark.parametrize("layout_nt", [False, True])
@pytest.mark.parametrize("n_step", [0, 1, 5, 200])
def test_bit_exact_vs_oracle(layout_nt, n_step):
    import torch

    from lerax_b200 import _lib as L
    from oracle import gae as G

    rng = np.random.default_rng(n_step)
    T, N, gamma = 77, 333, 0.97
    r, v = rng.normal(size=(T, N)).astype(f32), rng.normal(size=(T, N)).astype(f32)
    d, lv = rng.random((T, N)) < 0.06, rng.normal(size=N).astype(f32)
    want = G.bootstrapped_return(r, v, d, lv, gamma) if n_step == 0 else G.n_step_return(r, v, d, lv, gamma, n_step)
    lay = (lambda x: x.T) if layout_nt else (lambda x: x)
    rt, vt, dt = _t(lay(r)), _t(lay(v)), _t(lay(d).astype(np.uint8))
    adv, ret = torch.empty_like(rt), torch.empty_like(rt)
    L.check(L.lib.lrx_returns(L.ptr(rt), L.ptr(vt), L.ptr(dt), L.ptr(_t(lv)), L.ptr(adv), L.ptr(ret), N, T, gamma, n_step,
                              L.LRX_LAYOUT_NT if layout_nt else L.LRX_LAYOUT_TN, L.stream_ptr()), "lrx_returns")
    torch.cuda.synchronize()
    assert np.array_equal(lay(adv.cpu().numpy()), want[0])
    assert np.array_equal(lay(ret.cpu().numpy()), want[1])


def test_empty_and_invalid():
    from lerax_b200 import _lib as L

    assert L.lib.lrx_returns(None, None, None, None, None, None, 0, 5, 0.9, 0, L.LRX_LAYOUT_TN, L.stream_ptr()) == 0
    assert L.lib.lrx_returns(None, None, None, None, None, None, 5, 0, 0.9, 3, L.LRX_LAYOUT_TN, L.stream_ptr()) == 0
    assert L.lib.lrx_returns(None, None, None, None, None, None, 5, 5, 0.9, 3, L.LRX_LAYOUT_TN, L.stream_ptr()) != 0
    assert L.lib.lrx_returns(None, None, None, None, None, None, 5, 5, 0.9, -1, L.LRX_LAYOUT_TN, L.stream_ptr()) != 0


def test_rollout_buffer_reference_constructor_keeps_gae_api():
    """tests/test_buffer.py:58-82 of the reference: keyword constructor with per-env [T] arrays, stateless policy,
    action masks carried through, returns == advantages + values."""
    import torch

    from lerax_b200.buffer import RolloutBuffer
    from oracle import gae as G

    masks = torch.tensor([[True, True], [True, False]])
    buffer = RolloutBuffer(observations=np.array([[0.0], [1.0]], f32), actions=np.array([0, 1]),
                           rewards=np.array([1.0, 1.0], f32), dones=np.array([False, True]),
                           log_probs=np.array([-0.1, -0.2], f32), values=np.array([0.5, 0.25], f32), states=None,
                           action_masks=masks)
    updated = buffer.compute_returns_and_advantages(last_value=0.0, gae_lambda=0.95, gamma=0.99)
    assert updated.states is None and updated.action_masks is masks and buffer.action_masks is masks
    adv, ret, val = updated.advantages.cpu().numpy(), updated.returns.cpu().numpy(), updated.values.cpu().numpy()
    assert adv.shape == (2,) and np.all(np.isfinite(adv)) and np.allclose(ret, adv + val)
    want = G.gae(np.array([1.0, 1.0], f32), np.array([0.5, 0.25], f32), np.array([False, True]), f32(0.0), 0.99, 0.95)
    assert np.allclose(adv, want[0], rtol=1e-6) and np.allclose(ret, want[1], rtol=1e-6)
    assert updated.observations.shape == (2, 1) and updated.actions.tolist() == [0, 1]
