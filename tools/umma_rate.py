"""Tensor-pipe cost per tcgen05.mma (bf16, K = 16) for small shapes, constant descriptors."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lerax_b200 import _lib as L
for M in (64, 128):
    for N in (8, 16, 32, 64):
        if M == 128 and N % 16: continue
        for ds in (1, 2, 4, 8):
            cyc = torch.zeros(2, dtype=torch.int64, device="cuda")
            L.check(L.lib.lrx_debug_umma_rate(M, N, ds, L.ptr(cyc), L.stream_ptr()))
            torch.cuda.synchronize()
            c = cyc.cpu().numpy() / 256.0
            print(f"M={M:3d} N={N:2d} accumulators={ds}: issue {c[0]:6.1f}  issue+complete {c[1]:6.1f} cyc/MMA")
