#!/usr/bin/env python
"""Headline benchmark: env-steps/s of one full PPO iteration (rollout + GAE + PPO update).

    python bench.py --gpus N --steps K --warmup W [--config c2]   # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...       # the reference's CPU path

A "step" is ONE PPO iteration over one synthetic batch: a T-step rollout of the per-GPU env
shard through the fused rollout kernel, the GAE pass, and num_epochs x num_batches minibatch
updates (loss fwd/bwd + clip + Adam).  The default workload is BASELINE.json configs[1] ("c2"):
CartPole, MLPActorCriticPolicy defaults (12 995 parameters), 65 536 envs x 128 steps per GPU,
10 epochs x 32 minibatches; with N GPUs every rank owns its own 65 536-env shard (weak
scaling) and the minibatch gradients are summed over NVLink peer memory.  `--config` selects
the other named configurations of BASELINE.json (c1 README defaults, c3 Pendulum 256-wide,
c4 Acrobot 65 536 envs TOTAL = strong scaling, c5 GAE + update sweep).

The reference (pure Python/JAX) cannot run in this image (no jax wheels, no network), so the
reference arm times the repo's NumPy restatement of the reference (oracle/, fp32; pinned to the
reference's own source by tests/golden/refexec_*.npz) on the host cores, on a bounded sample of
the same workload -- see DESIGN.md "Measurement".
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "env-steps/sec (rollout+GAE+PPO update)"
UNIT = "env-steps/s"
UPDATE_KERNEL = "ppo_grad_tc128c_kernel"  # the tensor-core gradient kernel default-width policies run (profiles/kernels.json key)

WIDE = dict(feature_width=256, value_width=256, action_width=256)
# name -> env, envs (per GPU for weak scaling, TOTAL for strong), horizon, epochs, minibatches, policy kwargs, scaling
CONFIGS = {
    "c1": dict(env="cartpole", envs=4, horizon=2048, epochs=10, batches=32, policy={}, scaling="weak",
               what="CartPole PPO, README defaults (4 envs x 2048 steps, 10 epochs x 32 minibatches of 256 rows)"),
    "c2": dict(env="cartpole", envs=65536, horizon=128, epochs=10, batches=32, policy={}, scaling="weak",
               what="CartPole PPO, MLPActorCriticPolicy defaults, 65536 envs x 128 steps per GPU, 10 epochs x 32 minibatches"),
    "c3": dict(env="pendulum", envs=16384, horizon=128, epochs=10, batches=32, policy=WIDE, scaling="weak",
               what="Pendulum PPO, Gaussian policy, 2-layer 256-wide MLPs, 16384 envs x 128 steps per GPU, 10 epochs x 32 minibatches"),
    "c4": dict(env="acrobot", envs=65536, horizon=128, epochs=10, batches=32, policy={}, scaling="strong",
               what="Acrobot PPO, MLPActorCriticPolicy defaults, 65536 envs TOTAL x 128 steps sharded over the GPUs, 10 epochs x 32 minibatches"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS) + ["c5"])
    ap.add_argument("--envs", type=int, default=None, help="override: envs per GPU (total for a strong-scaling config)")
    ap.add_argument("--horizon", type=int, default=None)
    ap.add_argument("--epochs", type=int, default=None)
    ap.add_argument("--batches", type=int, default=None)
    ap.add_argument("--cpu-envs", type=int, default=4096, help="env count of the bounded CPU sample (65536 = the whole c2 workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.config != "c5":
        cfg = CONFIGS[args.config]
        for k in ("envs", "horizon", "epochs", "batches"):
            if getattr(args, k) is None:
                setattr(args, k, cfg[k])
        args.env, args.policy, args.scaling, args.workload = cfg["env"], cfg["policy"], cfg["scaling"], cfg["what"]
    return args


# ----------------------------------------------------------------------------- helpers
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], tflops=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    tflops_burst=p["bf16_tflops"], source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tflops=1400.0, tflops_burst=1590.0, source="fallback (B200_PROFILING.md)")


def profiled(kernel_key):
    """ncu-derived per-launch facts (dram bytes, executed flops) committed under profiles/kernels.json, keyed by
    '<kernel>|<env>|<envs>x<horizon>|<batches>'; None when this shape was never captured."""
    path = os.path.join(ROOT, "profiles", "kernels.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f).get(kernel_key)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def policy_flops_per_row(dims):
    """(fwd, dW, dX) MACs of the three Linear chains; dX skips the first encoder layer."""
    fwd = dw = dx = 0
    for c, chain in enumerate(dims):
        for i in range(len(chain) - 1):
            mac = chain[i] * chain[i + 1]
            fwd += mac
            dw += mac
            if not (c == 0 and i == 0):
                dx += mac
    return fwd, dw, dx


# -------------------------------------------------------------------- CPU (reference) arm
def cpu_threads_use_all():
    """Give the BLAS / OpenMP pools behind NumPy every host core (torchrun exports OMP_NUM_THREADS=1 to its ranks)."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits

        threadpool_limits(limits=n)
        got = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        got = 1
    return max(got, 1)


def cpu_iteration(env_name, pol_kw, n_envs, T, epochs, batches, seed=0):
    """One PPO iteration of the NumPy restatement of the reference; returns seconds."""
    import numpy as np

    from oracle import envs as OE, gae as OG, policy as OP, ppo as OU, rollout as OR

    oenv = {"cartpole": OE.cartpole, "pendulum": OE.pendulum, "acrobot": OE.acrobot}[env_name]()
    n_out = 1 if oenv.is_box else oenv.n_actions
    spec = OP.make_spec(oenv.obs_dim, n_out, oenv.is_box, **pol_kw)
    params = OP.init_params(spec, seed=seed)
    rng = np.random.default_rng(seed + 1)
    k = oenv.state_dim
    y = OE.initial_from_uniform(oenv, rng.random((k, n_envs)).astype(np.float32))
    na = (rng.standard_normal((T, n_envs)) if oenv.is_box else rng.gumbel(size=(T, n_envs, n_out))).astype(np.float32)
    ru = rng.random((T, n_envs, k)).astype(np.float32)
    idx = np.stack([OU.batch_indices(n_envs * T, n_envs * T // batches, rng) for _ in range(epochs)])
    t0 = time.perf_counter()
    ro = OR.rollout(oenv, spec, params, y, np.zeros(n_envs, np.float32), np.zeros(n_envs, np.int32), T, 0.99, na, ru)
    adv, ret = OG.gae(ro["rew"], ro["val"], ro["done"], ro["last_value"], 0.99, 0.95)
    flat = {k2: OU.flatten_time_major(v) for k2, v in
            dict(obs=ro["obs"], act=ro["act"], logp=ro["logp"], val=ro["val"], ret=ret, adv=adv).items()}
    OU.train(spec, params, OU.AdamState.init(params), flat, idx, OU.Hyper())
    return time.perf_counter() - t0


def cpu_sample_envs(args):
    return max(min(args.cpu_envs, args.envs), min(args.envs, 4))


def cpu_baseline(args, iters=1):
    cores = cpu_threads_use_all()
    n = cpu_sample_envs(args)
    cpu_iteration(args.env, args.policy, min(n, 256), args.horizon, 1, args.batches)  # warm caches / BLAS threads
    secs = [cpu_iteration(args.env, args.policy, n, args.horizon, args.epochs, args.batches, seed=i) for i in range(iters)]
    best = min(secs)
    return {
        "value": n * args.horizon / best, "unit": UNIT, "cores": cores, "kind": "port",
        "same_config": n == args.envs,
        "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"),
        "sample": f"{n} envs x {args.horizon} steps, {args.epochs} epochs x {args.batches} minibatches "
                  f"(1/{max(args.envs // n, 1)} of the per-GPU workload), NumPy fp32 oracle, {best:.2f} s/iteration; "
                  f"host has {os.cpu_count()} logical cores; real Lerax/JAX is not installable in this image",
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = cpu_threads_use_all()
    n = cpu_sample_envs(args)
    for _ in range(min(args.warmup, 1)):
        cpu_iteration(args.env, args.policy, min(n, 256), args.horizon, 1, args.batches)
    secs = [cpu_iteration(args.env, args.policy, n, args.horizon, args.epochs, args.batches, seed=i) for i in range(args.steps)]
    ms = 1e3 * sum(secs) / len(secs)
    value = n * args.horizon / (ms / 1e3)
    cb = {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "same_config": n == args.envs,
          "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"),
          "sample": f"{n} envs x {args.horizon} steps per step (bounded sample of the {args.envs}-env workload), "
                    f"NumPy fp32 restatement of the reference (JAX not installable offline)"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "name": args.config, "envs_per_gpu": args.envs, "horizon": args.horizon,
                   "epochs": args.epochs, "minibatches": args.batches, "sample_envs": n},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------ our arm
def make_env(name):
    from lerax_b200.env.classic_control import Acrobot, CartPole, Pendulum

    return {"cartpole": CartPole, "pendulum": Pendulum, "acrobot": Acrobot}[name]()


def parity_check(algo, state, buf, world, dist):
    """Multi-GPU only: from identical state, three minibatch updates through the NVLink peer-memory exchange and through
    the NCCL all-reduce fallback must agree, and every rank must hold bit-identical parameters afterwards."""
    import torch

    from lerax_b200 import random as jr

    def three_updates(disable_p2p):
        pol = state.policy.with_params(state.policy.params.clone())
        opt = state.opt_state.clone()
        a = type(algo)(num_envs=algo.num_envs, num_steps=algo.num_steps, num_epochs=1, num_batches=algo.num_batches)
        a._bufs, a._force_nccl = algo._bufs, disable_p2p
        if not disable_p2p and getattr(algo, "_ex", None) is not None:
            a._ex, a._ex_key = algo._ex, algo._ex_key  # keep the sequence numbers of the live exchange
        a.minibatches_per_epoch = 3
        a.train(pol, opt, buf, key=jr.key(77))
        torch.cuda.synchronize()
        return pol.params

    p_p2p = three_updates(False)
    p_nccl = three_updates(True)
    rel = ((p_p2p - p_nccl).abs() / (p_nccl.abs() + 1e-6)).max()
    gathered = [torch.empty_like(p_p2p) for _ in range(world)]
    dist.all_gather(gathered, p_p2p)
    bit_equal = all(torch.equal(gathered[0], g) for g in gathered[1:])
    flags = torch.tensor([float(rel), 0.0 if bit_equal else 1.0], dtype=torch.float64, device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MAX)
    moved = float((p_p2p - state.policy.params).abs().max())
    return {"ranks_bit_equal": flags[1].item() == 0.0, "p2p_vs_nccl_max_rel": flags[0].item(), "minibatches": 3,
            "max_param_step": moved, "exchange": "nvlink-p2p" if getattr(algo, "_ex", None) is not None else "nccl"}


def grad_launch_alone_us(algo, state, buf, world):
    """The dominant kernel by itself: one epoch's minibatches through `lrx_ppo_grad` only (gradient kernel + partial
    reduction into gradbuf; no exchange, no clip, no Adam -- the parameters do not move), CUDA events on the launching
    stream.  Returns (us per launch pair, launches)."""
    import torch

    from lerax_b200 import _lib as L
    from lerax_b200 import random as jr

    b = algo._bufs
    view = buf.view(packed=b.get("packed"))
    B_global = algo.batch_size
    B_local = B_global // world
    idx = buf.batch_indices(B_local, key=jr.key(12345))[: algo.minibatches_per_epoch]
    nb = idx.shape[0]
    L.check(L.lib.lrx_ppo_adv_sums(view.adv, view.N, view.T, L.ptr(idx), nb, B_local, L.ptr(b["adv_sums"]), L.stream_ptr()),
            "lrx_ppo_adv_sums")
    pol_d, h = state.policy.desc(), algo._hyper(algo._lr_at(0))
    ws, ws_bytes = L.ptr(b["workspace"]), b["workspace"].numel()

    def epoch():
        for i in range(nb):
            L.check(L.lib.lrx_ppo_grad(pol_d, L.ptr(state.policy.params), view, L.ptr(idx[i]), B_local, B_global,
                                       L.ptr(b["adv_sums"][i]), h, L.ptr(b["gradbuf"]), ws, ws_bytes, L.stream_ptr()),
                    "lrx_ppo_grad")

    epoch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    epoch()
    epoch()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / (2 * nb), 2 * nb


def run_ours(args):
    import torch
    import torch.distributed as dist

    from lerax_b200 import _lib as L
    from lerax_b200 import random as jr
    from lerax_b200.algorithm import PPO
    from lerax_b200.buffer import RolloutBuffer
    from lerax_b200.policy import MLPActorCriticPolicy

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; lerax_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    env = make_env(args.env)
    strong = args.scaling == "strong"
    n_total = args.envs if strong else args.envs * world
    n_local = n_total // world
    algo = PPO(num_envs=n_total, num_steps=args.horizon, num_epochs=args.epochs, num_batches=args.batches)
    policy = MLPActorCriticPolicy(env, key=jr.key(0), **args.policy)
    cb = algo.consolidate_callbacks(None)
    state = algo.reset(env, policy, key=jr.key(1), callback=cb)
    keys = jr.split(jr.key(2), args.warmup + args.steps)
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: W warm-up iterations, then exactly K timed ones
    for k in keys[: args.warmup]:
        state = algo.iteration(state, key=k, callback=cb)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    L.lib.lrx_launch_count(1)
    barrier()
    buf = None
    for i, k in enumerate(keys[args.warmup:]):
        rollout_key, train_key, _ = jr.split(k, 3)
        ev[i][0].record(stream)
        b = algo._buffers(env, state.policy, state.policy.params.device)
        out = L.LrxRolloutOut(L.ptr(b["obs"]), L.ptr(b["act"]), L.ptr(b["rew"]), L.ptr(b["done"]), L.ptr(b["logp"]),
                              L.ptr(b["val"]), L.ptr(b["last_value"]), None, None, None, L.ptr(b["rollout_ws"]),
                              0 if b["rollout_ws"] is None else b["rollout_ws"].numel())
        rng = L.LrxRng(state.rollout_seed, rank * n_local, state.steps_done & 0xFFFFFFFF)
        L.check(L.lib.lrx_rollout(env.desc(), state.policy.desc(), L.ptr(state.policy.params),
                                  state.step_state.env_state.c_state(), rng, None, out, n_local, args.horizon,
                                  float(algo.gamma), L.stream_ptr()), "lrx_rollout")
        ev[i][1].record(stream)
        buf = RolloutBuffer(b["obs"], b["act"], b["rew"], b["done"], b["logp"], b["val"])
        buf = buf.compute_returns_and_advantages(b["last_value"], algo.gae_lambda, algo.gamma, ev_sums=b["ev_sums"])
        ev[i][2].record(stream)
        applied = state.iteration_count * algo.num_epochs * algo.minibatches_per_epoch
        algo.train(state.policy, state.opt_state, buf, key=train_key, applied=applied)
        state.iteration_count += 1
        state.steps_done += args.horizon
        ev[i][3].record(stream)
    end = torch.cuda.Event(enable_timing=True)
    end.record(stream)
    barrier()
    launches = int(L.lib.lrx_launch_count(0))
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0][0].elapsed_time(end)
    stage_ms = [sum(e[j].elapsed_time(e[j + 1]) for e in ev) / args.steps for j in range(3)]
    t = torch.tensor([total_ms] + stage_ms, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, roll_ms, gae_ms, upd_ms = t.tolist()
    algo.check_finite()
    ms_per_step = total_ms / args.steps
    steps_per_iter = n_total * args.horizon
    value = steps_per_iter / (ms_per_step / 1e3)

    par = parity_check(algo, state, buf, world, dist) if world > 1 else None
    grad_us = grad_launch_alone_us(algo, state, buf, world) if rank == 0 and algo.minibatches_per_epoch > 0 else None
    if world > 1:
        dist.barrier()

    # ---- end to end through the public API with HOST parameters: learn() for one iteration
    e2e = None
    if not args.no_e2e:
        host_params = torch.from_numpy(policy.numpy()).pin_memory()
        host_out = torch.empty_like(host_params).pin_memory()
        e2e_algo = PPO(num_envs=n_total, num_steps=args.horizon, num_epochs=args.epochs, num_batches=args.batches)

        def e2e_step(i):
            pol = policy.with_params(host_params)  # parameters start on the HOST (pinned)
            trained = e2e_algo.learn(env, pol, total_timesteps=steps_per_iter, key=jr.key(100 + i))
            host_out.copy_(trained.params, non_blocking=True)  # D2H of the result
            return trained

        for i in range(max(1, min(args.warmup, 2))):
            e2e_step(i)
        barrier()
        s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_ev.record(stream)
        for i in range(args.steps):
            e2e_step(10 + i)
        e_ev.record(stream)
        barrier()
        t2 = torch.tensor([s_ev.elapsed_time(e_ev)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e_ms = t2.item() / args.steps
        pbytes = host_params.numel() * 4
        e2e = {"value": steps_per_iter / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": pbytes, "d2h_bytes_per_step": pbytes + 4,
               "api": "PPO(...).learn(env, policy_with_host_params, total_timesteps=N*T, key=...) -> params to host"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        if par is not None and not (par["ranks_bit_equal"] and par["p2p_vs_nccl_max_rel"] < 1e-4):
            raise SystemExit(3)
        return

    peaks = measured_peaks()
    fwd, dw, dx = policy_flops_per_row(policy._dims)
    rows = n_local * args.horizon  # per GPU
    n_mb = args.epochs * algo.minibatches_per_epoch
    mb_rows = algo.batch_size // world
    upd_flops = 2.0 * (fwd + dw + dx) * mb_rows * n_mb
    roll_flops = 2.0 * fwd * rows + 300.0 * rows
    row_bytes = 4 * policy.obs_dim + 4 + 4 + 1 + 4 + 4
    stages = {
        "rollout": {"ms": roll_ms, "env_steps_per_s": rows / (roll_ms / 1e3), "tflops": roll_flops / (roll_ms / 1e3) / 1e12,
                    "hbm_gbs": rows * row_bytes / (roll_ms / 1e3) / 1e9, "bytes_per_env_step": row_bytes},
        "gae": {"ms": gae_ms, "hbm_gbs": rows * 17 / (gae_ms / 1e3) / 1e9, "frac_of_hbm_peak": rows * 17 / (gae_ms / 1e3) / 1e9 / peaks["hbm_gbs"],
                "bytes_per_element": 17, "note": "event-timed inside the iteration, right after the rollout wrote the same planes "
                "(partly L2-resident); cold-L2 numbers per shape: profiles/r2_gae.md"},
        "update": {"ms": upd_ms, "tflops": upd_flops / (upd_ms / 1e3) / 1e12, "flop_per_row_epoch": 2 * (fwd + dw + dx),
                   "us_per_minibatch": 1e3 * upd_ms / n_mb, "minibatch_rows_per_gpu": mb_rows},
    }
    dominant = max(("rollout", roll_ms), ("gae", gae_ms), ("update", upd_ms), key=lambda x: x[1])[0]
    default_widths = not args.policy
    kernel_name = {"update": UPDATE_KERNEL if default_widths else "generic layer-by-layer update",
                   "rollout": "rollout_tc_kernel" if default_widths else "rollout_generic_kernel", "gae": "gae_tn_kernel"}[dominant]
    prof = profiled(f"{kernel_name}|{args.env}|{n_local}x{args.horizon}|{args.batches}")
    if dominant == "gae":
        roof = {"bound": "hbm", "achieved": stages["gae"]["hbm_gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s"}
    else:
        roof = {"bound": "tensor", "achieved": stages[dominant]["tflops"], "peak": peaks["tflops"], "unit": "TFLOP/s"}
    roof.update({"frac": roof["achieved"] / roof["peak"], "traffic": prof.get("dram_bytes_per_launch") if prof else None,
                 "kernel": kernel_name, "peak_source": peaks["source"],
                 "note": "achieved = ALGORITHMIC fp32 FLOPs (2 x (fwd + dW + dX) MACs per row-epoch; rollout: 2 x fwd MACs + ~300 "
                         "for the Tsit5 step) / CUDA-event time of the stage over the timed region; peak = sustained bf16 "
                         "tensor throughput. `traffic`, `executed_*` come from the committed ncu capture of this kernel and "
                         "shape (profiles/kernels.json) and are null when that shape was never captured"})
    if dominant == "update" and grad_us is not None:
        roof.update({"kernel_us_alone": grad_us[0], "kernel_launches_alone": grad_us[1],
                     "achieved_kernel_alone": 2.0 * (fwd + dw + dx) * mb_rows / (grad_us[0] * 1e-6) / 1e12,
                     "frac_kernel_alone": 2.0 * (fwd + dw + dx) * mb_rows / (grad_us[0] * 1e-6) / 1e12 / peaks["tflops"],
                     "kernel_alone_note": "lrx_ppo_grad only (gradient kernel(s) + partial reduction), 2 epochs of minibatches "
                                          "back to back after the timed region; `achieved` above divides by the whole update "
                                          "stage (also advantage sums, permutation, clip + Adam)"})
    if prof and prof.get("executed_flop_per_row_epoch") and dominant == "update":
        executed = prof["executed_flop_per_row_epoch"] * mb_rows * n_mb / (upd_ms / 1e3) / 1e12
        roof.update({"executed_tflops": executed, "executed_frac": executed / peaks["tflops"], "profile": prof.get("source")})

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "name": args.config, "envs_per_gpu": n_local, "envs_total": n_total,
                   "horizon": args.horizon, "epochs": args.epochs, "minibatches": args.batches, "params": policy.n_params,
                   "l2": "rollout buffer (%.0f MB per GPU) %s the 126 MB L2; no explicit flush" % (
                       rows * (row_bytes + 8) / 1e6, "exceeds" if rows * (row_bytes + 8) > 126e6 else "FITS in"),
                   "parallelism": f"env-sharded x{world}, per-minibatch gradient sum over NVLink peer memory" if world > 1 else "single GPU"},
        "clocks": clocks, "gpu_launches": launches, "launches_per_step": launches / args.steps, "stages": stages, "roofline": roof,
    }
    if par is not None:
        line["parity_check"] = par
    if e2e is not None:
        line["e2e"] = e2e
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if par is not None and not (par["ranks_bit_equal"] and par["p2p_vs_nccl_max_rel"] < 1e-4):
        raise SystemExit(3)


# ------------------------------------------------------------------- c5: GAE + update sweep
def run_sweep(args):
    """BASELINE.json configs[4]: GAE and one-epoch PPO update microbenchmarks, T in {128, 512} x N in {4k .. 1M}, next to
    the NumPy oracle on a bounded sample (rank 0, one GPU)."""
    import numpy as np
    import torch

    from lerax_b200 import _lib as L
    from lerax_b200 import random as jr
    from lerax_b200.algorithm import PPO
    from lerax_b200.buffer import RolloutBuffer
    from lerax_b200.policy import MLPActorCriticPolicy
    from oracle import gae as OG

    torch.cuda.set_device(0)
    peaks = measured_peaks()
    env = make_env("cartpole")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    rows_out = []
    for T in (128, 512):
        for N in (4096, 16384, 65536, 262144, 1048576):
            if N * T > (1 << 29):
                continue
            r = torch.randn(T, N, device="cuda"); v = torch.randn(T, N, device="cuda")
            d = (torch.rand(T, N, device="cuda") < 0.02).to(torch.uint8); lv = torch.randn(N, device="cuda")
            adv = torch.empty_like(r); ret = torch.empty_like(r)
            evs = torch.zeros(L.LRX_GAE_EV_DOUBLES, dtype=torch.float64, device="cuda")
            f = lambda: L.check(L.lib.lrx_gae(L.ptr(r), L.ptr(v), L.ptr(d), L.ptr(lv), L.ptr(adv), L.ptr(ret), N, T, 0.99, 0.95,
                                              0, L.ptr(evs), L.stream_ptr()))
            for _ in range(3):
                f()
            tot = 0.0
            for _ in range(5):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); f(); e1.record(); torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            gae_ms = tot / 5
            # CPU oracle on a bounded strip of the same arrays
            nc = min(N, 4096)
            rc, vc, dc = (x[:, :nc].cpu().numpy() for x in (r, v, d))
            lc = lv[:nc].cpu().numpy()
            t0 = time.perf_counter(); OG.gae(rc, vc, dc.astype(bool), lc, 0.99, 0.95); cpu_s = time.perf_counter() - t0
            del r, v, d, adv, ret
            # one epoch of PPO updates (32 minibatches) on a rollout produced by the kernel itself
            upd_ms = None
            if N * T <= (1 << 27):
                algo = PPO(num_envs=N, num_steps=T, num_epochs=1, num_batches=32)
                pol = MLPActorCriticPolicy(env, key=jr.key(0))
                cb = algo.consolidate_callbacks(None)
                st = algo.reset(env, pol, key=jr.key(1), callback=cb)
                for k in jr.split(jr.key(2), 2):
                    st = algo.iteration(st, key=k, callback=cb)
                buf = st.last_buffer
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(3):
                    algo.train(st.policy, st.opt_state, buf, key=jr.key(10 + i))
                e1.record(); torch.cuda.synchronize()
                upd_ms = e0.elapsed_time(e1) / 3
                algo._bufs = None
                del algo, st, buf
                torch.cuda.empty_cache()
            fwd, dw, dx = policy_flops_per_row([(4, 64, 64, 16), (16, 64, 64, 1), (16, 64, 16, 2)])
            row = {"T": T, "N": N, "gae_us": 1e3 * gae_ms, "gae_gbs": N * T * 17 / gae_ms / 1e6,
                   "gae_frac_hbm": N * T * 17 / gae_ms / 1e6 / peaks["hbm_gbs"],
                   "gae_cpu_elems_per_s": nc * T / cpu_s, "gae_gpu_elems_per_s": N * T / (gae_ms / 1e3)}
            if upd_ms is not None:
                row.update({"update_epoch_ms": upd_ms, "update_rows_per_s": N * T / (upd_ms / 1e3),
                            "update_tflops": 2.0 * (fwd + dw + dx) * N * T / (upd_ms / 1e3) / 1e12})
            rows_out.append(row)
            print(json.dumps(row), file=sys.stderr, flush=True)
    print(json.dumps({"metric": "GAE GB/s and PPO-update rows/s sweep", "config": {"workload": "c5: GAE (cold L2) + one-epoch PPO "
                      "update (32 minibatches, CartPole default policy) sweep", "name": "c5"}, "unit": "GB/s | rows/s",
                      "n_gpus": 1, "data": "synthetic", "dtype": "f32", "sweep": rows_out,
                      "value": max(r["gae_gbs"] for r in rows_out), "higher_is_better": True}))


def main():
    args = parse()
    if args.config == "c5":
        if int(os.environ.get("RANK", "0")) == 0:
            run_sweep(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
