"""Time line of the first issue steps of ppo_grad_pp64_kernel (CTA 0, half 0) at the C2 minibatch shape: raw clock64
stamps -> cycles relative to the step's request.  Run on a B200."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lerax_b200 import _lib as L
from lerax_b200 import random as jr
from lerax_b200.algorithm import PPO
from lerax_b200.env.classic_control import CartPole
from lerax_b200.policy import MLPActorCriticPolicy

dbg = L.debug_lib()
dbg.lrx_debug_set_fused(3)
env = CartPole()
algo = PPO(num_envs=65536, num_steps=128, num_epochs=1, num_batches=32)
pol = MLPActorCriticPolicy(env, key=jr.key(0))
cb = algo.consolidate_callbacks(None)
st = algo.reset(env, pol, key=jr.key(1), callback=cb)
st = algo.iteration(st, key=jr.key(2), callback=cb)
torch.cuda.synchronize()
R = 12 * 8
prof = torch.zeros(R * 8, dtype=torch.int64, device="cuda")
dbg.lrx_debug_set_pp64_prof(prof.data_ptr(), R)
algo.minibatches_per_epoch = 1
algo.train(st.policy, st.opt_state, st.last_buffer, key=jr.key(3))
torch.cuda.synchronize()
dbg.lrx_debug_set_pp64_prof(None, 0)
p = prof.cpu().view(R, 8).tolist()
names = ["enc1", "enc2", "v0", "v1", "v1bwd", "a0", "a1", "a1bwd", "a0bwd", "enc2bwd", "enc1bwd", "dW0"]
print("step        req->arrive  ->issuer saw  ->committed  ->crit done  ->tmem ld  ->math  ->stored(next req)   total")
for r in range(3 * 12, 6 * 12):
    t = p[r]
    base = t[0]
    f = lambda i: (t[i] - base) if t[i] else -1
    nxt = p[r + 1][0] - base if r + 1 < R else -1
    print(f"{names[r % 12]:8s} {f(1):12d} {f(2):12d} {f(3):12d} {f(4):12d} {f(5):10d} {f(6):8d} {f(7):12d} {nxt:10d}")
