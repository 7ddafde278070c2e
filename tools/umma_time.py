"""Cycles per tcgen05.mma (kind::f16, bf16) for the shapes / operand views of the fused kernels.
dsplit = number of independent TMEM accumulators the consecutive MMAs rotate over."""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lerax_b200 import _lib as L

def kmaj(MN): return [0, 128, MN * 16, MN * 16, 128, 2 * MN * 16]
def mnmaj(kr): return [1, kr * 16, 128, 128, kr * 16, 256]

def t(M, N, K, ca, cb, nprod=6, reps=64, dsplit=1):
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda")
    dump = torch.zeros(128, N, device="cuda"); cyc = torch.zeros(2, dtype=torch.int64, device="cuda")
    arr = (C.c_int32 * 12)(*(ca + cb))
    L.check(L.debug_lib().lrx_debug_umma_time_bf16(L.ptr(A), L.ptr(B), L.ptr(dump), M, N, K, arr, nprod | (dsplit << 8), reps, L.ptr(cyc), L.stream_ptr()))
    torch.cuda.synchronize()
    n = reps * (K // 16) * nprod
    c = cyc.cpu().numpy()
    return c[0] / n, c[1] / n

for M in (64, 128):
    for N in (8, 16, 32, 64):
        if M == 128 and N % 16: continue
        for ds in (1, 2, 4):
            if ds * N > 256: continue
            for name, ca, cb in (("K/K", kmaj(M), kmaj(N)), ("MN/MN", mnmaj(64), mnmaj(64))):
                i, tot = t(M, N, 64, ca, cb, dsplit=ds)
                print(f"M={M:3d} N={N:3d} dsplit={ds} {name:5s}: issue {i:6.1f}, issue+complete {tot:6.1f} cyc/MMA")
