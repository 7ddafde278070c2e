"""The specialised tensor-core kernels (default policy widths) against the generic-width kernels of the
same library, and size-independent properties at BASELINE.json's full sizes (65 536 envs x 128 steps):
determinism, GAE returns = advantages + values, the update's gradient linearity over disjoint row sets."""

import numpy as np
import pytest

from oracle import gae as OG
from oracle import ppo as OU

import helpers as H
from test_gpu_parity import noise_for, setup, synthetic_buffer

pytestmark = pytest.mark.gpu
KINDS = ["cartpole", "pendulum", "acrobot", "mountain_car", "continuous_mountain_car"]


@pytest.fixture
def fused_toggle():
    from lerax_b200 import _lib as L

    yield L.debug_lib().lrx_debug_set_fused  # 4: 128-row kernel with composite weights (default), 3: two half tiles in flight, 2: 128-row tensor-core kernels, 1: first-generation update kernel, 0: generic
    L.debug_lib().lrx_debug_set_fused(4)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("mes", [0, 9])
def test_rollout_tensor_core_vs_generic(kind, mes, fused_toggle):
    N, T = 200, 48  # two CTAs of the 128-env kernel, the second one ragged
    oenv, spec, params, env, pol, rng, y0 = setup(kind, N=N, mes=mes)
    na, ru = noise_for(spec, rng, T, N, oenv.state_dim)
    t0, sc0 = np.zeros(N, np.float32), np.zeros(N, np.int32)
    fused_toggle(2)
    a = H.run_rollout(env, pol, y0, t0, sc0, T, 0.99, na, ru)
    fused_toggle(0)
    b = H.run_rollout(env, pol, y0, t0, sc0, T, 0.99, na, ru)
    assert (a["done"] != b["done"]).mean() == 0
    if not spec.is_box:
        assert (a["act"] != b["act"]).mean() == 0
    # per step from identical states: value / log-prob of the two implementations at 1e-5
    same = np.isclose(a["y_trace"], b["y_trace"], rtol=0, atol=0).all(axis=1)  # [T, N] bit-identical inputs
    assert same[0].all()
    for name in ("val", "logp", "rew"):
        H.assert_close(a[name][same], b[name][same], 1e-5, 2e-6, name + " (identical states)")
    drift = 1e-5 if kind == "cartpole" else 3e-4
    for name in ("obs", "val", "logp", "last_value"):
        H.assert_close(a[name], b[name], 1e-4, drift, name)


@pytest.mark.parametrize("level", [4, 3, 2, 1])
@pytest.mark.parametrize("kind", KINDS)
def test_gradient_tensor_core_vs_generic(kind, level, fused_toggle):
    T, N, B = 40, 33, 1000  # 7 full 128-row tiles (15 of 64 rows for the first-generation kernel) + a ragged one
    spec, params, pol, rng, flat, buf = synthetic_buffer(kind, T, N, seed=21)
    hp = OU.Hyper(clip_value_loss=True, entropy_loss_coefficient=0.01)
    idx = rng.permutation(T * N)[:B].astype(np.int32)
    fused_toggle(level)
    ga, _ = H.run_grad(pol, buf, idx, hp)
    ga2, _ = H.run_grad(pol, buf, idx, hp)
    np.testing.assert_array_equal(ga, ga2)  # deterministic (fixed-order reductions)
    fused_toggle(0)
    gb, _ = H.run_grad(pol, buf, idx, hp)
    P = spec.n_params
    scale = np.abs(gb[:P]).max()
    H.assert_close(ga[:P], gb[:P], 1e-4, 1e-5 * scale, "gradient")
    H.assert_close(ga[P:P + 4], gb[P:P + 4], 1e-5, 1e-6, "stat sums")


def test_full_size_properties():
    """BASELINE configs[1]: 65 536 envs x 128 steps.  Too large for the oracle; checked through
    size-independent properties."""
    import torch

    from lerax_b200 import _lib as L

    N, T = 65536, 128
    oenv, spec, params, env, pol, rng, _ = setup("cartpole", N=8)
    y0 = rng.uniform(-0.05, 0.05, (4, N)).astype(np.float32)
    t0, sc0 = np.zeros(N, np.float32), np.zeros(N, np.int32)
    a = H.run_rollout(env, pol, y0, t0, sc0, T, 0.99, seed=11, trace=False)
    b = H.run_rollout(env, pol, y0, t0, sc0, T, 0.99, seed=11, trace=False)
    for k in ("obs", "act", "rew", "done", "logp", "val", "last_value"):
        np.testing.assert_array_equal(a[k], b[k])  # bit-reproducible
    assert np.isfinite(a["val"]).all() and np.isfinite(a["logp"]).all() and (a["logp"] <= 0).all()
    assert set(np.unique(a["act"])) <= {0, 1} and (a["rew"] == 1.0).all()
    ep_len = T * N / max(a["done"].sum(), 1)
    assert 8 < ep_len < 60  # a random CartPole policy falls after ~10-40 steps
    # the first 512 envs reproduce the oracle's GAE; everywhere returns = advantages + values
    adv, ret = H.run_gae(a["rew"], a["val"], a["done"], a["last_value"], 0.99, 0.95)
    oadv, oret = OG.gae(a["rew"][:, :512], a["val"][:, :512], a["done"][:, :512], a["last_value"][:512], 0.99, 0.95)
    H.assert_close(adv[:, :512], oadv, 1e-5, 1e-5, "adv")
    np.testing.assert_allclose(ret, adv + a["val"], rtol=0, atol=1e-5)
    # gradient of a 262 144-row minibatch = sum of the gradients of its two halves (each scaled by
    # B_global): linearity of the mean loss over disjoint row sets
    planes = dict(obs=a["obs"], act=a["act"], logp=a["logp"], val=a["val"], ret=ret, adv=adv)
    buf = H.DeviceBuffer(**planes)
    hp = OU.Hyper(normalize_advantages=False)
    idx = rng.permutation(T * N)[:262144].astype(np.int32)
    g_all, _ = H.run_grad(pol, buf, idx, hp)
    g_1, _ = H.run_grad(pol, buf, idx[:131072], hp, B_global=262144)
    g_2, _ = H.run_grad(pol, buf, idx[131072:], hp, B_global=262144)
    P = spec.n_params
    scale = np.abs(g_all[:P]).max()
    H.assert_close(g_1[:P] + g_2[:P], g_all[:P], 1e-4, 2e-6 * scale, "gradient linearity")
    g_again, _ = H.run_grad(pol, buf, idx, hp)
    np.testing.assert_array_equal(g_all, g_again)  # 28 tiles per CTA, prefetched rows, fixed-order reductions
    torch.cuda.synchronize()


@pytest.mark.parametrize("level", [4, 3, 2, 1])
@pytest.mark.parametrize("kind", ["cartpole", "pendulum", "mountain_car"])
def test_gradient_many_tiles_per_cta_vs_generic(kind, level, fused_toggle):
    """Every CTA of the fused kernel walks two or three tiles (row staging reuse, buffer reuse across tiles, a ragged
    last tile); the generic layer-by-layer kernels are the independent implementation."""
    rows = 128 if level >= 2 else 64  # level 3: two 64-row half tiles per CTA and round
    T, N = 64, 400 * rows // 64
    B = rows * 148 * 2 + 37
    spec, params, pol, rng, flat, buf = synthetic_buffer(kind, T, N, seed=5)
    hp = OU.Hyper(clip_value_loss=True, entropy_loss_coefficient=0.01)
    idx = rng.permutation(T * N)[:B].astype(np.int32)
    fused_toggle(level)
    ga, _ = H.run_grad(pol, buf, idx, hp)
    ga2, _ = H.run_grad(pol, buf, idx, hp)
    np.testing.assert_array_equal(ga, ga2)
    fused_toggle(0)
    gb, _ = H.run_grad(pol, buf, idx, hp)
    P = spec.n_params
    scale = np.abs(gb[:P]).max()
    H.assert_close(ga[:P], gb[:P], 1e-4, 1e-5 * scale, "gradient")
    H.assert_close(ga[P:P + 4], gb[P:P + 4], 1e-5, 1e-6, "stat sums")


@pytest.mark.parametrize("n", [1, 2, 7, 256, 1000, 65536, 8388608])
def test_permutation_is_a_bijection(n):
    import torch

    from lerax_b200 import _lib as L

    out = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    L.check(L.lib.lrx_permutation(L.ptr(out), n, 0x1234567890ABCDEF, L.stream_ptr()), "lrx_permutation")
    assert torch.equal(torch.sort(out).values, torch.arange(n, dtype=torch.int32, device="cuda"))
    if n >= 256:
        out2 = torch.empty_like(out)
        L.check(L.lib.lrx_permutation(L.ptr(out2), n, 0x1234567890ABCDF0, L.stream_ptr()), "lrx_permutation")
        assert (out != out2).float().mean() > 0.9  # another key, another permutation
        assert (out == torch.arange(n, dtype=torch.int32, device="cuda")).float().mean() < 0.05
        # no structure a minibatch would notice: first-quarter mean index close to n/2
        q = out[: n // 4].double().mean().item()
        assert abs(q - (n - 1) / 2) < 0.05 * n


@pytest.mark.parametrize("kind", ["cartpole", "acrobot", "pendulum"])
def test_packed_rows_gather_equals_plane_gather(kind):
    """LrxBatchView.packed (lrx_ppo_pack_rows: one 64-byte record per row, reference order) vs the gather from the
    [T][N] planes: same rows, same arithmetic -> bit-identical gradients."""
    import torch

    from lerax_b200 import _lib as L

    T, N, B = 48, 50, 2000
    spec, params, pol, rng, flat, buf = synthetic_buffer(kind, T, N, seed=17)
    # the record of row i = n*T + t holds exactly the planes' values at (t, n)
    rec = buf.packed.cpu().numpy().reshape(N, T, L.LRX_PACKED_FLOATS)
    np.testing.assert_array_equal(rec[:, :, :spec.obs_dim], np.swapaxes(buf.t["obs"].cpu().numpy(), 0, 1))
    np.testing.assert_array_equal(rec[:, :, spec.obs_dim:8], 0)
    for k, name in ((9, "logp"), (10, "val"), (11, "ret"), (12, "adv")):
        np.testing.assert_array_equal(rec[:, :, k], buf.t[name].cpu().numpy().T)
    np.testing.assert_array_equal(rec[:, :, 8].view(np.int32), buf.t["act"].cpu().numpy().T.view(np.int32))
    hp = OU.Hyper(clip_value_loss=True, entropy_loss_coefficient=0.01)
    idx = rng.permutation(T * N)[:B].astype(np.int32)
    ga, _ = H.run_grad(pol, buf, idx, hp)
    buf.view.packed = None
    gb, _ = H.run_grad(pol, buf, idx, hp)
    np.testing.assert_array_equal(ga, gb)
    torch.cuda.synchronize()


@pytest.mark.parametrize("widths", [dict(feature_width=256, value_width=256, action_width=256),
                                    dict(feature_width=96, value_width=40, action_width=72, feature_size=24, value_depth=3)])
def test_wide_policy_gradient_tensor_core_gemm_vs_fma(widths, fused_toggle):
    """Non-default widths (BASELINE configs[2]: 256-wide): the layer-by-layer update with the tcgen05 GEMM
    (lrx_gemm_tc.cu: all four operand orientations, ragged tiles, split-K weight gradients) against the same path on the
    FMA kernel, and both against the oracle."""
    T, N, B = 40, 60, 2100  # 16 full 128-row tiles + a ragged one; split-K over 3 row chunks
    spec, params, pol, rng, flat, buf = synthetic_buffer("pendulum", T, N, seed=29, **widths)
    hp = OU.Hyper(clip_value_loss=True, entropy_loss_coefficient=0.01)
    idx = rng.permutation(T * N)[:B].astype(np.int32)
    fused_toggle(2)
    ga, _ = H.run_grad(pol, buf, idx, hp)
    ga2, _ = H.run_grad(pol, buf, idx, hp)
    np.testing.assert_array_equal(ga, ga2)
    fused_toggle(0)
    gb, _ = H.run_grad(pol, buf, idx, hp)
    P = spec.n_params
    scale = np.abs(gb[:P]).max()
    H.assert_close(ga[:P], gb[:P], 1e-4, 1e-5 * scale, "gradient (tensor-core GEMM vs FMA)")
    H.assert_close(ga[P:P + 4], gb[P:P + 4], 1e-5, 1e-6, "stat sums")
    _, stats, g = OU.ppo_loss_and_grad(spec, params, OU.gather(flat, idx), hp)
    gg = ga[:P].copy()
    gg[-1] -= hp.entropy_loss_coefficient
    H.assert_close(gg, g, 1e-4, 1e-5 * np.abs(g).max(), "gradient vs oracle")


@pytest.mark.parametrize("kind,mes", [("pendulum", 0), ("pendulum", 6), ("cartpole", 5), ("acrobot", 0)])
def test_wide_policy_stepwise_rollout(kind, mes, fused_toggle):
    """Non-default widths, >= 512 envs: the step-by-step GEMM-batched rollout (lrx_rollout_steps.cu) against the oracle
    and against the one-lane-per-env kernel -- identical event order (PPO.step: sample, clip, transition, reward,
    terminal / truncate, TimeLimit bootstrap from the PRE-reset state, reset), same injected noise."""
    widths = dict(feature_width=96, value_width=80, action_width=72, feature_size=24)
    N, T = 600, 20
    oenv, spec, params, env, pol, rng, y0 = setup(kind, N=N, mes=mes, **widths)
    na, ru = noise_for(spec, rng, T, N, oenv.state_dim)
    t0, sc0 = np.zeros(N, np.float32), np.zeros(N, np.int32)
    from oracle import rollout as OR

    want = OR.rollout(oenv, spec, params, y0, t0, sc0, T, 0.99, na, ru)
    fused_toggle(2)
    a = H.run_rollout(env, pol, y0, t0, sc0, T, 0.99, na, ru)           # stepwise (workspace given, N >= 512)
    b = H.run_rollout(env, pol, y0, t0, sc0, T, 0.99, na, ru, workspace=False)  # one lane per env
    for got, name in ((a, "stepwise"), (b, "per-lane")):
        assert (got["done"] != want["done"]).mean() == 0, name
        if not spec.is_box:
            assert (got["act"] != want["act"]).mean() == 0, name
        for k in ("obs", "logp", "val", "rew"):
            H.assert_close(got[k][0], want[k][0], 1e-5, 2e-6, f"{name} {k}[t=0]")
        drift = 1e-5 if kind == "cartpole" else 3e-4
        for k in ("obs", "logp", "val", "rew", "last_value", "y"):
            H.assert_close(got[k], want[k], 1e-4, drift, f"{name} {k}")
        np.testing.assert_array_equal(got["step_count"], want["step_count"])
    if mes:
        assert want["done"].any()
    # Philox noise (no replay): both kernels draw the same streams
    c = H.run_rollout(env, pol, y0, t0, sc0, T, 0.99, seed=5, trace=False)
    d = H.run_rollout(env, pol, y0, t0, sc0, T, 0.99, seed=5, trace=False, workspace=False)
    assert (c["done"] != d["done"]).mean() < 2e-3
    H.assert_close(c["val"][0], d["val"][0], 1e-5, 2e-6, "Philox: value[t=0]")
