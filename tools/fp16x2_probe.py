"""Accuracy of an fp16x2 split (x = h0 + h1, 3 products) vs the bf16x3 split (6 products) against float64, for
operand magnitudes like the PPO path's (weights ~0.1, activations ~1, gradients small).  Run on a B200."""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lerax_b200 import _lib as L

def kmaj(MN): return [0, 128, MN * 16, MN * 16, 128, 2 * MN * 16]

def run(A, B, nprod):
    M, K = A.shape; N = B.shape[0]
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    dump = torch.zeros(128, N, device="cuda")
    arr = (C.c_int32 * 12)(*(kmaj(M) + kmaj(N)))
    L.check(L.debug_lib().lrx_debug_umma_gemm_bf16(L.ptr(dA), L.ptr(dB), L.ptr(dump), M, N, K, arr, nprod, L.stream_ptr()))
    torch.cuda.synchronize()
    return dump.cpu().numpy()

rng = np.random.default_rng(0)
for sa, sb, tag in [(1.0, 1.0, "a~1, b~1"), (1.0, 0.1, "act~1, w~0.1"), (1e-3, 0.1, "act~1e-3, w~0.1"), (1e-5, 0.1, "grad~1e-5, w~0.1"),
                    (100.0, 0.1, "act~100, w~0.1"), (3e4, 1.0, "act~3e4")]:
    A = (rng.standard_normal((128, 64)) * sa).astype(np.float32); B = (rng.standard_normal((64, 64)) * sb).astype(np.float32)
    want = A.astype(np.float64) @ B.astype(np.float64).T
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
    f32 = (A @ B.T).astype(np.float64)
    e_b = np.abs(run(A, B, 6) - want) / scale
    e_h = np.abs(run(A, B, 3 | 0x10000) - want) / scale
    e_f = np.abs(f32 - want) / scale
    print(f"{tag:22s}: bf16x3 {e_b.max():.2e}   fp16x2 {e_h.max():.2e}   numpy fp32 {e_f.max():.2e}   (max err / sum|a||b|)")
