"""Per-step cycle profile of ppo_grad_tc128_kernel (CTA 0) at the C2 minibatch shape (run on a B200)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lerax_b200 import _lib as L
from lerax_b200 import random as jr
from lerax_b200.algorithm import PPO
from lerax_b200.env.classic_control import CartPole
from lerax_b200.policy import MLPActorCriticPolicy

N, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (65536, 128)
env = CartPole()
algo = PPO(num_envs=N, num_steps=T, num_epochs=1, num_batches=32)
pol = MLPActorCriticPolicy(env, key=jr.key(0))
cb = algo.consolidate_callbacks(None)
st = algo.reset(env, pol, key=jr.key(1), callback=cb)
st = algo.iteration(st, key=jr.key(2), callback=cb)
torch.cuda.synchronize()
prof = torch.zeros(160, dtype=torch.int64, device="cuda")
L.debug_lib().lrx_debug_set_tc128_prof(prof.data_ptr())
algo.train(st.policy, st.opt_state, st.last_buffer, key=jr.key(3))
torch.cuda.synchronize()
L.debug_lib().lrx_debug_set_tc128_prof(None)
p = prof.cpu().tolist()
launches = 32
tiles = (N * T // 32 + 127) // 128
tiles_cta0 = (tiles + 147) // 148
level = int(os.environ.get("LRX_FUSED", "4"))
names4 = ["", "enc0 (MMA + epi)", "enc1 (MMA + epi)", "heads layer 0 (composite)", "v1 + value dot", "value loss + dY",
          "v1 bwd + a0 fwd epi", "logits + policy loss", "a1 bwd", "h2 bwd (composite)", "enc1 bwd", "g0 only step", "top: wait rows + XOBS", "-",
          "(loop overhead)", "tail wait", "flush"]
names = ["", "top: enc0 FMA + XH1", "enc1 (MMA + epi)", "enc2", "v0 (+a0 bg)", "v1 + value dot", "value loss + dY", "v1 bwd + a0 fwd epi",
         "a1 + policy loss", "a1 bwd", "a0 bwd -> dfeat", "enc2 bwd", "enc1 bwd", "g0 only step", "(loop overhead)", "tail wait", "flush"]
if level >= 4:
    names = names4
tot = sum(p[1:15])
print(f"{tiles_cta0} tiles per CTA, {launches} launches; cycles per tile (row-owner thread 0 | issuer lane 0)")
for k in range(1, 17):
    per = launches * tiles_cta0 if k <= 14 else launches
    print(f"{k:2d} {names[k]:28s} {p[k] / per:9.0f} | {p[32 + k] / per:9.0f}")
print("barrier wait per step (thread 0 | issuer):")
for k in range(1, 12):
    per = launches * tiles_cta0
    if level < 4:
        print(f"  step {k:2d}: {p[16 + k] / per:7.0f} | {p[48 + k] / per:7.0f}")
print(f"tile total {tot / (launches * tiles_cta0):.0f} cycles")
print("flush per launch (thread 0): barrier + TMEM -> staging + register sums %d | barrier %d | staging -> global + final sums %d" % (
    p[64] / launches, p[65] / launches, p[66] / launches))
print("flush phases (thread 0 | issuer):", [round(p[64 + k] / launches) for k in range(7)], [round(p[96 + k] / launches) for k in range(7)])
