import sys, os, torch
os.environ["LRX_PROBE_ST_SHAPE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lerax_b200 import _lib as L
d = torch.zeros(128, 8, device="cuda")
L.check(L.debug_lib().lrx_debug_tmem_shape_probe(L.ptr(d), L.stream_ptr())); torch.cuda.synchronize()
d = d.cpu().numpy().astype(int)
print("TMEM (lane, col 0..7) <- (thread, reg) for tcgen05.st.16x128b.x1  [-1 = untouched]")
for lane in list(range(0, 18)) + [31]:
    print(lane, [(x // 2, x % 2) if x >= 0 else None for x in d[lane]])
