"""The generic-width update's GEMM (lrx_gemm_tc.cu) by itself, at the shapes of the 256-wide policy of BASELINE.json
configs[2]: forward, dX (with ReLU mask) and split-K weight gradient, for a range of minibatch rows, CUDA events, L2
flushed before every launch.  Prints us per launch, the algorithmic bytes it moves and GB/s, next to the FMA kernel.
Usage (GPU box): python tools/gemm_time.py [width]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lerax_b200 import _lib as L

W = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lib = L.debug_lib()


def timed(fn, reps=6):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return 1e3 * tot / reps


def gemm(A, lda, B, ldb, Cm, ldc, M, N, K, at, bt, bias=None, relu=0, mask=None, ldm=0, splits=1, k_chunk=0, stride=0, tc=1):
    L.check(lib.lrx_debug_gemm(L.ptr(A), lda, L.ptr(B), ldb, L.ptr(Cm), ldc, M, N, K, at, bt, L.ptr(bias) if bias is not None else None,
                               relu, L.ptr(mask) if mask is not None else None, ldm, 0, splits, k_chunk, stride, tc, L.stream_ptr()),
            "lrx_debug_gemm")


print(f"width {W}; us per launch (tcgen05 | FMA), algorithmic MB, GB/s of the tcgen05 kernel")
for rows in (8192, 16384, 32768, 65536, 131072):
    Wt = torch.randn(W, W, device=dev) / 16
    bias = torch.randn(W, device=dev)
    X = torch.randn(W + 1, rows, device=dev)
    Y = torch.empty(W + 1, rows, device=dev)
    dY = torch.randn(W, rows, device=dev)
    thin = torch.randn(17, rows, device=dev)
    Wthin = torch.randn(W, 16, device=dev)
    splits = max(1, rows // 1024)
    part = torch.empty(splits, W * (W + 1) + 64, device=dev)
    cases = {
        "fwd  W x W": (lambda tc: gemm(Wt, W, X, rows, Y, rows, W, rows, W, 0, 0, bias=bias, relu=1, tc=tc), 2 * W * rows * 4),
        "fwd 16 -> W": (lambda tc: gemm(Wthin, 16, thin, rows, Y, rows, W, rows, 16, 0, 0, bias=bias, relu=1, tc=tc), W * rows * 4),
        "dX   W x W": (lambda tc: gemm(Wt, W, dY, rows, Y, rows, W, rows, W, 1, 0, mask=X, ldm=rows, tc=tc), 3 * W * rows * 4),
        "dW   W x W": (lambda tc: gemm(dY, rows, X, rows, part, W, W, W + 1, rows, 0, 1, splits=splits, k_chunk=1024,
                                        stride=part.shape[1], tc=tc), 2 * W * rows * 4),
    }
    for name, (fn, nbytes) in cases.items():
        t_tc = timed(lambda: fn(1))
        t_fma = timed(lambda: fn(0), reps=2)
        print(f"rows {rows:7d}  {name:12s} {t_tc:8.1f} | {t_fma:8.1f} us   {nbytes / 1e6:7.1f} MB  {nbytes / t_tc / 1e3:7.0f} GB/s")
