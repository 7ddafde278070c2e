"""One cold-L2 GAE launch per shape (for `ncu -k regex:gae_tn`): T x N from argv, default the C2 shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lerax_b200 import _lib as L

T, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 65536)
r = torch.randn(T, N, device="cuda"); v = torch.randn(T, N, device="cuda")
d = (torch.rand(T, N, device="cuda") < 0.02).to(torch.uint8); lv = torch.randn(N, device="cuda")
adv = torch.empty_like(r); ret = torch.empty_like(r)
ev = torch.zeros(L.LRX_GAE_EV_DOUBLES, dtype=torch.float64, device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.fill_(1)
    L.check(L.lib.lrx_gae(L.ptr(r), L.ptr(v), L.ptr(d), L.ptr(lv), L.ptr(adv), L.ptr(ret), N, T, 0.99, 0.95, 0, L.ptr(ev), L.stream_ptr()))
torch.cuda.synchronize()
print("ev sums", ev[:4].tolist())
