"""Where a minibatch of the PPO update goes: one lrx_ppo_update epoch vs the same minibatches through lrx_ppo_grad
only (fused gradient kernel + partial reduction, no Adam), timed with CUDA events on a bench-sized buffer.
Usage (GPU box): python tools/update_breakdown.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lerax_b200 import _lib as L
from lerax_b200 import random as jr
from lerax_b200.algorithm import PPO
from lerax_b200.env.classic_control import CartPole
from lerax_b200.policy import MLPActorCriticPolicy

env = CartPole()
algo = PPO(num_envs=65536, num_steps=128, num_epochs=1, num_batches=32)
policy = MLPActorCriticPolicy(env, key=jr.key(0))
cb = algo.consolidate_callbacks(None)
state = algo.reset(env, policy, key=jr.key(1), callback=cb)
state = algo.iteration(state, key=jr.key(2), callback=cb)  # fills the buffers, warms everything
b = algo._bufs
from lerax_b200.buffer import RolloutBuffer
buf = RolloutBuffer(b["obs"], b["act"], b["rew"], b["done"], b["logp"], b["val"])
buf = buf.compute_returns_and_advantages(b["last_value"], algo.gae_lambda, algo.gamma, ev_sums=b["ev_sums"])
view = buf.view(packed=b.get("packed"))
B = algo.batch_size
idx = buf.batch_indices(B, key=jr.key(3))[:32]
nb = idx.shape[0]
pol = state.policy
opt = state.opt_state
d = pol.desc()
h = algo._hyper(algo._lr_at(0))
ws, wsb = L.ptr(b["workspace"]), b["workspace"].numel()
stats = b["stats"]

def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def epoch_update():
    L.check(L.lib.lrx_ppo_update(d, L.ptr(pol.params), L.ptr(opt.mu), L.ptr(opt.nu), L.ptr(opt.count), view, L.ptr(idx),
                                 nb, B, h, L.ptr(stats[0]), L.ptr(b["nonfinite"]), ws, wsb, L.stream_ptr()), "update")

def epoch_grad_only():
    for i in range(nb):
        L.check(L.lib.lrx_ppo_grad(d, L.ptr(pol.params), view, L.ptr(idx[i]), B, B, L.ptr(b["adv_sums"][i]), h,
                                   L.ptr(b["gradbuf"]), ws, wsb, L.stream_ptr()), "grad")

def epoch_grad_apply():
    for i in range(nb):
        L.check(L.lib.lrx_ppo_grad(d, L.ptr(pol.params), view, L.ptr(idx[i]), B, B, L.ptr(b["adv_sums"][i]), h,
                                   L.ptr(b["gradbuf"]), ws, wsb, L.stream_ptr()), "grad")
        L.check(L.lib.lrx_ppo_apply(d, L.ptr(pol.params), L.ptr(opt.mu), L.ptr(opt.nu), L.ptr(opt.count),
                                    L.ptr(b["gradbuf"]), h, L.ptr(stats[0, i]), L.ptr(b["nonfinite"]), L.stream_ptr()), "apply")

# the data-parallel path with a one-rank exchange (every NVLink access is local): what the p2p kernels cost by themselves
import ctypes as C
from lerax_b200._p2p import GradientExchange
own = C.c_void_p()
hbuf = (C.c_ubyte * L.LRX_P2P_HANDLE_BYTES)()
L.check(L.lib.lrx_p2p_alloc(pol.n_params + L.LRX_PPO_NSTATS, C.byref(own), hbuf), "alloc")
ex = GradientExchange(pol.n_params + L.LRX_PPO_NSTATS, 0, 1, own.value, [], torch.tensor([own.value], dtype=torch.int64, device="cuda"),
                      torch.zeros((), dtype=torch.int32, device="cuda"))

def epoch_p2p_self():
    ex.update_epoch(d, pol, opt, view, idx, B, B, b["adv_sums"], h, None, stats[0], b["nonfinite"], b["workspace"])

L.check(L.lib.lrx_ppo_adv_sums(view.adv, view.N, view.T, L.ptr(idx), nb, B, L.ptr(b["adv_sums"]), L.stream_ptr()), "sums")
t_u = timed(epoch_update)
t_g = timed(epoch_grad_only)
t_ga = timed(epoch_grad_apply)
t_p = timed(epoch_p2p_self)
print(f"per minibatch (B={B}): lrx_ppo_update {1e3 * t_u / nb:.1f} us | lrx_ppo_grad only {1e3 * t_g / nb:.1f} us | "
      f"grad + apply (NCCL path without the all-reduce) {1e3 * t_ga / nb:.1f} us | "
      f"lrx_ppo_update_p2p with a one-rank exchange {1e3 * t_p / nb:.1f} us")
