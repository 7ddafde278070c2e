"""Decode K-major addressing for every descriptor layout_type (run on a B200)."""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lerax_b200 import _lib as L
np.set_printoptions(linewidth=250)

def run(M, N, K, cfg, A, B, split=1):
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    dump = torch.full((128, N), float("nan"), device="cuda")
    arr = (C.c_int32 * 12)(*cfg)
    L.check(L.debug_lib().lrx_debug_umma_gemm(L.ptr(dA), L.ptr(dB), L.ptr(dump), M, N, K, arr, split, L.stream_ptr()))
    torch.cuda.synchronize()
    return dump.cpu().numpy()

def kmaj(MN):
    return [0, 128, MN * 16, MN * 16, 128, 2 * MN * 16]

M, N, K = 128, 64, 8
eye = np.zeros((M, K), np.float32); eye[np.arange(8), np.arange(8)] = 1
Bz = np.zeros((N, K), np.float32)
for lt in (0, 1, 2, 4, 6):
    for lbo, sbo in [(16, 1024), (2048, 1024), (16, 512)]:
        d = run(M, N, K, kmaj(M) + [2 | (lt << 4), 0, 0, lbo, sbo, 0], eye, Bz)
        w = d[:8, :].astype(np.int64) * 4   # w[k][n]
        print(f"B K-major raw lt={lt} lbo={lbo} sbo={sbo}: nnz={np.count_nonzero(d)}")
        for n in (0, 1, 2, 3, 4, 7, 8, 9, 16):
            print(f"      n={n}: k0..7 -> {list(w[:, n])}")
# A-side MN-major lt=1 decode (M = 128): B = identity K-major N = 16
M, N, K = 128, 16, 8
Bi = np.zeros((N, K), np.float32); Bi[np.arange(8), np.arange(8)] = 1
Az = np.zeros((M, K), np.float32)
for lbo, sbo in [(4096, 512), (512, 4096)]:
    d = run(M, N, K, [3 | (1 << 4), 0, 0, lbo, sbo, 0] + kmaj(N), Az, Bi)
    w = d[:, :8].astype(np.int64) * 4   # w[m][k]
    print(f"A MN-major raw lt=1 lbo={lbo} sbo={sbo}: nnz={np.count_nonzero(d)}")
    for m in (0, 1, 8, 31, 32, 33, 64, 96):
        print(f"      m={m}: k0..7 -> {list(w[m, :])}")
