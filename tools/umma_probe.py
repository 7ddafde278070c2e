"""Decode how tcgen05.mma addresses an MN-major shared-memory operand (run on a B200).
The probed operand's region is filled with its own word index; the other operand is an
identity / ones, so the accumulator shows which words the hardware read for each (mn, k)."""
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

def kmaj(MN):  # canonical K-major for one K=8 MMA: chunk-major with R = MN rows
    return [0, 128, MN * 16, MN * 16, 128, 2 * MN * 16]

M, N, K = 128, 64, 8
ones = np.ones((M, K), np.float32)
eye = np.zeros((M, K), np.float32); eye[np.arange(8), np.arange(8)] = 1
Bz = np.zeros((N, K), np.float32)
for lt in (0, 1, 2, 4, 6):
    for lbo, sbo in [(128, 1024), (1024, 128)]:
        d = run(M, N, K, kmaj(M) + [3 | (lt << 4), 0, 0, lbo, sbo, 0], eye, Bz)
        w = d[:8, :].astype(np.int64) * 4
        print(f"B mn raw lt={lt} lbo={lbo} sbo={sbo}: nnz={np.count_nonzero(d)} nan={np.isnan(d).sum()}  k=0: n0..9={list(w[0,:10])} k=1: {list(w[1,:6])} n=32:{w[0,32]}")
# sanity: the same raw-fill probe on a K-major B (known good)
d = run(M, N, K, kmaj(M) + [2, 0, 0, N * 16, 128, 0], eye, Bz)
w = d[:8, :].astype(np.int64) * 4
print(f"B K-major raw: nnz={np.count_nonzero(d)}  k=0: n0..9={list(w[0,:10])} k=0..7@n=0: {list(w[:,0])}")
# real data, b MN-major in the canonical storage, report error vs expectation and vs zero
rng = np.random.default_rng(0)
A = rng.standard_normal((M, K)).astype(np.float32); B = rng.standard_normal((N, K)).astype(np.float32)
want = A.astype(np.float64) @ B.astype(np.float64).T
for lbo, sbo, smn, sk in [(128, K*16, K*16, 128), (K*16, 128, K*16, 128), (128, 128, 128, N//4*128), (N//4*128, 128, 128, N//4*128)]:
    d = run(M, N, K, kmaj(M) + [1, smn, sk, lbo, sbo, 128], A, B)
    print(f"real b_mn storage(smn={smn},sk={sk}) lbo={lbo} sbo={sbo}: max|d-want|={np.abs(d-want).max():.3e}  max|d|={np.abs(d).max():.3e}")
