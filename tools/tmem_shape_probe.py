import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lerax_b200 import _lib as L
d = torch.zeros(128, 8, device="cuda")
L.check(L.debug_lib().lrx_debug_tmem_shape_probe(L.ptr(d), L.stream_ptr())); torch.cuda.synchronize()
d = d.cpu().numpy().astype(int)
for t in list(range(0, 12)) + [31, 32, 33]:
    print(t, [(x // 100, x % 100) for x in d[t]])
