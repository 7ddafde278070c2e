"""GAE kernel time vs steps-per-thread chunk (run on a B200)."""
import sys, os, torch, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lerax_b200 import _lib as L
L.lib.lrx_debug_set_gae_chunk.argtypes = [C.c_int]
for T, N in ((128, 65536), (512, 65536), (128, 1 << 20)):
    r = torch.randn(T, N, device="cuda"); v = torch.randn(T, N, device="cuda")
    d = (torch.rand(T, N, device="cuda") < 0.02).to(torch.uint8); lv = torch.randn(N, device="cuda")
    adv = torch.empty_like(r); ret = torch.empty_like(r); ev = torch.zeros(4, dtype=torch.float64, device="cuda")
    for chunk in (16, 18, 8, 24):  # 16/18: one env per lane, 16/8 steps; 8/24: two envs per lane, 8/4 steps
        L.lib.lrx_debug_set_gae_chunk(chunk)
        for with_ev in (True, False):
            f = lambda: L.check(L.lib.lrx_gae(L.ptr(r), L.ptr(v), L.ptr(d), L.ptr(lv), L.ptr(adv), L.ptr(ret), N, T, 0.99, 0.95, 0, L.ptr(ev) if with_ev else None, L.stream_ptr()))
            for _ in range(3): f()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(20): f()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f"T={T} N={N} chunk={chunk} ev={with_ev}: {ms*1000:7.1f} us  {T*N*17/ms/1e6:7.0f} GB/s")
