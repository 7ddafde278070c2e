"""Tensor-core building blocks (lerax_b200/csrc/lrx_umma.cuh, lrx_tc.cuh) against float64 NumPy:
every operand orientation and MMA shape the fused rollout / update kernels use -- the bf16x3 fp32
emulation in the canonical shared-memory layout (K-major and MN-major views of the same bytes),
the TMEM lane mapping of M = 64 accumulators, and the 3xTF32 K-major form."""

import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def kmaj16(MN):  # 2-byte canonical layout of an MN-row matrix, K-major view (K = 16 per MMA)
    return [0, 128, MN * 16, MN * 16, 128, 2 * MN * 16]


def mnmaj16(krows):  # same bytes, MN-major view: the stored matrix has `krows` rows (= K)
    return [1, krows * 16, 128, 128, krows * 16, 256]


def kmaj32(MN):  # 4-byte (tf32) canonical K-major layout (K = 8 per MMA)
    return [0, 128, MN * 16, MN * 16, 128, 2 * MN * 16]


def _run(fn, M, N, K, cfg, flag, seed=0):
    import torch
    from lerax_b200 import _lib as L

    rng = np.random.default_rng(seed)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    dump = torch.full((128, N), float("nan"), device="cuda")
    arr = (C.c_int32 * 12)(*cfg)
    L.check(getattr(L.debug_lib(), fn)(L.ptr(dA), L.ptr(dB), L.ptr(dump), M, N, K, arr, flag, L.stream_ptr()), fn)
    torch.cuda.synchronize()
    d = dump.cpu().numpy()
    if M == 64:  # row m lives in TMEM lane (m % 16) + 32 * (m // 16)
        m = np.arange(64)
        d = d[(m % 16) + 32 * (m // 16)]
    want = A.astype(np.float64) @ B.astype(np.float64).T
    scale = np.abs(A.astype(np.float64)) @ np.abs(B.astype(np.float64)).T
    return np.abs(d - want) / scale


BF16_CASES = [
    # M, N, K, A view, B view                      role in the fused kernels
    (128, 64, 64, "k", "k"),   # rollout forward 64->64 (128 envs per CTA)
    (128, 16, 64, "k", "k"),   # forward 64->16
    (128, 64, 16, "k", "k"),   # forward 16->64
    (64, 64, 64, "k", "k"),    # update forward, 64-row tile
    (64, 16, 64, "k", "k"),
    (64, 64, 64, "k", "mn"),   # dX = dY W  (W read MN-major)
    (64, 16, 64, "k", "mn"),
    (64, 64, 16, "k", "mn"),
    (64, 64, 64, "mn", "mn"),  # dW^T = dY^T X  (contraction over the 64 rows)
    (64, 16, 64, "mn", "mn"),
    (64, 8, 64, "mn", "mn"),
]


@pytest.mark.parametrize("M,N,K,va,vb", BF16_CASES)
def test_bf16x3_matches_fp64(M, N, K, va, vb):
    ca = kmaj16(M) if va == "k" else mnmaj16(K)
    cb = kmaj16(N) if vb == "k" else mnmaj16(K)
    err = _run("lrx_debug_umma_gemm_bf16", M, N, K, ca + cb, 6)
    assert err.max() < 1.5e-6, f"bf16x3 error {err.max():.3e} of sum|a||b|"


def test_bf16_fewer_products_are_less_accurate():
    e6 = _run("lrx_debug_umma_gemm_bf16", 128, 64, 64, kmaj16(128) + kmaj16(64), 6).max()
    e3 = _run("lrx_debug_umma_gemm_bf16", 128, 64, 64, kmaj16(128) + kmaj16(64), 3).max()
    e1 = _run("lrx_debug_umma_gemm_bf16", 128, 64, 64, kmaj16(128) + kmaj16(64), 1).max()
    assert e6 < 1.5e-6 < e3 < 1e-4 < e1 < 2e-2


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (128, 16, 64), (128, 64, 16), (128, 64, 8), (64, 64, 64)])
def test_3xtf32_k_major_matches_fp64(M, N, K):
    err = _run("lrx_debug_umma_gemm", M, N, K, kmaj32(M) + kmaj32(N), 1)
    assert err.max() < 2e-6, f"3xTF32 error {err.max():.3e}"
    err1 = _run("lrx_debug_umma_gemm", M, N, K, kmaj32(M) + kmaj32(N), 0)
    assert 1e-6 < err1.max() < 2e-3, f"plain TF32 error {err1.max():.3e}"
