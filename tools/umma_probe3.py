"""bf16 (kind::f16) operand addressing and bf16x3 accuracy (run on a B200)."""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lerax_b200 import _lib as L
np.set_printoptions(linewidth=250)

def run(M, N, K, cfg, A, B, nprod=6):
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    dump = torch.full((128, N), float("nan"), device="cuda")
    arr = (C.c_int32 * 12)(*cfg)
    L.check(L.debug_lib().lrx_debug_umma_gemm_bf16(L.ptr(dA), L.ptr(dB), L.ptr(dump), M, N, K, arr, nprod, L.stream_ptr()))
    torch.cuda.synchronize()
    return dump.cpu().numpy()

def kmaj(MN):   # chunk-major canonical, K-major view: s_mn, s_k, lbo, sbo, kstep
    return [0, 128, MN * 16, MN * 16, 128, 2 * MN * 16]
def mnmaj(Krows, lbo_is_k=True):  # same bytes viewed MN-major (stored matrix has Krows rows)
    return [1, Krows * 16, 128] + ([128, Krows * 16] if lbo_is_k else [Krows * 16, 128]) + [256]

rng = np.random.default_rng(0)
def check(M, N, K, ca, cb, tag):
    A = rng.standard_normal((M, K)).astype(np.float32); B = rng.standard_normal((N, K)).astype(np.float32)
    want = A.astype(np.float64) @ B.astype(np.float64).T
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
    for nprod in (6, 3, 1):
        d = run(M, N, K, ca + cb, A, B, nprod)
        got = d if M == 128 else d[(np.arange(64) % 16) + 32 * (np.arange(64) // 16)]
        print(f"{tag} M={M} N={N} K={K} nprod={nprod}: max rel err {np.abs(got - want).max() / 1:.3e} (scaled {(np.abs(got-want)/scale).max():.3e}) max|d|={np.abs(got).max():.2f}")

check(128, 64, 64, kmaj(128), kmaj(64), "K/K")
check(128, 16, 16, kmaj(128), kmaj(16), "K/K")
# raw decode of MN-major B for several layout types
M, N, K = 128, 64, 16
eye = np.zeros((M, K), np.float32); eye[np.arange(16), np.arange(16)] = 1
Bz = np.zeros((N, K), np.float32)
for lt in (0, 2, 4, 6):
    for lbo, sbo in [(128, 1024), (1024, 128)]:
        try:
            d = run(M, N, K, kmaj(M) + [3 | (lt << 4), 0, 0, lbo, sbo, 0], eye, Bz)
        except Exception as e:
            print("lt", lt, lbo, sbo, "FAULT", str(e)[:60]); raise
        w = d[:16, :].astype(np.int64) * 2   # byte offsets, w[k][n]
        print(f"B MN raw lt={lt} lbo={lbo} sbo={sbo}: nnz={np.count_nonzero(d)} k=0: n0..10={list(w[0,:11])} | n=0: k0..15={list(w[:,0])} | n=8,16,64@k0: {w[0,8]},{w[0,16]}")
check(128, 64, 64, kmaj(128), mnmaj(64, True), "K/MN(lbo=k)")
check(128, 64, 64, kmaj(128), mnmaj(64, False), "K/MN(lbo=mn)")
check(64, 72, 64, mnmaj(64, True), mnmaj(64, True), "MN/MN(lbo=k)")
check(64, 72, 64, mnmaj(64, False), mnmaj(64, False), "MN/MN(lbo=mn)")
