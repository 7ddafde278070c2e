"""Does a tcgen05.ld wait behind MMAs already queued on the tensor pipe?  Latency of eight loads from an idle TMEM
region with 0 / 16 / 64 / 256 MMAs (M = 64, N = 64: ~32 clk each) accumulating into another region; then the same
loading the accumulator itself WITHOUT waiting for the commit (is the value the finished sum?)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lerax_b200 import _lib as L
for n in (0, 16, 256, 10016, 10256, -16, -256):
    cyc = torch.zeros(16, dtype=torch.int64, device="cuda")
    for _ in range(2):
        L.check(L.debug_lib().lrx_debug_tmem_ld_under_mma(n, L.ptr(cyc), L.stream_ptr()))
    torch.cuda.synchronize()
    c = cyc.cpu().numpy()
    if n >= 10000:
        print("first MMA accumulates onto uninitialised TMEM:", end=" ")
        n -= 10000
    what = "other region" if n >= 0 else f"accumulator, no commit wait: read {c[9]} (finished sum = {16 * -n})"
    print(f"{abs(n):4d} MMAs queued, {what}: ld latencies {[int(x) for x in c[:8]]}  loads + commit wait {c[8]} cycles")
