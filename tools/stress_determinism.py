"""Race hunt: 60 repeats of the fused PPO gradient on the same inputs must be bit-identical (three env kinds,
up to 28 tiles per CTA).  compute-sanitizer is not available on the pool, so determinism + agreement with the
independent generic kernels is the race check.  Usage (GPU box): python tools/stress_determinism.py"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers as H
from test_gpu_parity import synthetic_buffer
from oracle import ppo as OU
for kind, T, N, B in (("cartpole", 128, 4096, 262144), ("pendulum", 64, 400, 64 * 148 * 2 + 37), ("acrobot", 64, 1000, 50000)):
    spec, params, pol, rng, flat, buf = synthetic_buffer(kind, T, N, seed=7)
    hp = OU.Hyper(clip_value_loss=True, entropy_loss_coefficient=0.01)
    idx = rng.permutation(T * N)[:B].astype(np.int32)
    ref, _ = H.run_grad(pol, buf, idx, hp)
    bad = 0
    for i in range(60):
        g, _ = H.run_grad(pol, buf, idx, hp)
        bad += int(not np.array_equal(g, ref))
    print(kind, "mismatching repeats:", bad, "finite:", bool(np.isfinite(ref).all()))
