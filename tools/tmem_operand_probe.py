import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lerax_b200 import _lib as L
np.set_printoptions(linewidth=220)
def run(mode, variant, M):
    d = torch.full((128, 16), float("nan"), device="cuda")
    L.check(L.debug_lib().lrx_debug_tmem_operand_probe(L.ptr(d), mode, variant, M, L.stream_ptr())); torch.cuda.synchronize()
    return d.cpu().numpy()
for M in (128, 64):
    d = run(0, 0, M)
    print(f"A-from-TMEM M={M}: accumulator lane -> source lane id (col 0):", d[:, 0].astype(int).tolist())
    print("   row 0 all cols (should be constant):", d[0].astype(int).tolist())
    d = run(0, 1, M)
    print(f"A-from-TMEM M={M}: k index seen at output column n (lane 0):", d[0].astype(int).tolist())
d = run(1, 0, 64)
print("M=64 accumulator at lane offset 16: lane -> row id:", d[:, 0].astype(int).tolist())
