"""Tensor-pipe cost per tcgen05.mma (bf16, K = 16) for small shapes, constant descriptors, for the four
combinations of K-major / MN-major operands (the weight-gradient GEMMs read both operands MN-major)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lerax_b200 import _lib as L
for a_mn, b_mn in ((0, 0), (0, 1), (1, 0), (1, 1)):
    for M in (64, 128):
        for N in (8, 16, 64):
            if M == 128 and N % 16: continue
            for ds in (1, 4):
                cyc = torch.zeros(2, dtype=torch.int64, device="cuda")
                L.check(L.debug_lib().lrx_debug_umma_rate(M, N, ds | (a_mn << 8) | (b_mn << 9), L.ptr(cyc), L.stream_ptr()))
                torch.cuda.synchronize()
                c = cyc.cpu().numpy() / 256.0
                print(f"A {'MN' if a_mn else 'K '}-major B {'MN' if b_mn else 'K '}-major M={M:3d} N={N:2d} accumulators={ds}: "
                      f"issue {c[0]:6.1f}  issue+complete {c[1]:6.1f} cyc/MMA")
