"""Tensor-pipe cost per tcgen05.mma (bf16, K = 16): A operand from shared memory (SS) vs from TMEM (TS), one or two
warps issuing concurrently into separate accumulators."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lerax_b200 import _lib as L
for M in (64, 128):
    for N in (16, 64):
        for ts in (0, 1):
            for issuers in (1, 2):
                cyc = torch.zeros(8, dtype=torch.int64, device="cuda")
                L.check(L.debug_lib().lrx_debug_umma_rate(M, N, 2 | (ts << 10) | ((issuers - 1) << 11), L.ptr(cyc), L.stream_ptr()))
                torch.cuda.synchronize()
                c = cyc.cpu().numpy() / 256.0
                print(f"{'TS' if ts else 'SS'} M={M:3d} N={N:2d} issuers={issuers}: per-warp issue {c[0]:6.1f} / {c[2]:6.1f}  issue+complete {c[1]:6.1f} / {c[3]:6.1f} cyc per MMA of ONE warp")
