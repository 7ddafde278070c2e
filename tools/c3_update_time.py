"""One epoch of the wide-policy (BASELINE.json configs[2]: Pendulum, 256-wide MLPs) PPO update, timed with CUDA events:
us per minibatch through `PPO.train` (layer-by-layer tcgen05 GEMM path).  Usage (GPU box): python tools/c3_update_time.py
[envs horizon batches]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lerax_b200 import random as jr
from lerax_b200.algorithm import PPO
from lerax_b200.env.classic_control import Pendulum
from lerax_b200.policy import MLPActorCriticPolicy

N, T, NB = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (16384, 128, 32)
env = Pendulum()
algo = PPO(num_envs=N, num_steps=T, num_epochs=1, num_batches=NB)
pol = MLPActorCriticPolicy(env, feature_width=256, value_width=256, action_width=256, key=jr.key(0))
cb = algo.consolidate_callbacks(None)
st = algo.reset(env, pol, key=jr.key(1), callback=cb)
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
torch.cuda.synchronize()
e[0].record()
st = algo.iteration(st, key=jr.key(2), callback=cb)
e[1].record()
torch.cuda.synchronize()
print(f"first iteration (rollout + GAE + 1 epoch): {e[0].elapsed_time(e[1]):.1f} ms")
buf = st.last_buffer
reps = 3
e[2].record()
for i in range(reps):
    algo.train(st.policy, st.opt_state, buf, key=jr.key(10 + i))
e[3].record()
torch.cuda.synchronize()
ms = e[2].elapsed_time(e[3]) / reps
print(f"update epoch: {ms:.2f} ms = {1e3 * ms / NB:.1f} us per minibatch of {N * T // NB} rows ({pol.n_params} params)")
