#!/usr/bin/env python
"""Headline benchmark: env-steps/s of one full PPO iteration (rollout + GAE + PPO update).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

A "step" is ONE PPO iteration over one synthetic batch: a T-step rollout of the per-GPU env
shard through the fused rollout kernel, the GAE pass, and num_epochs x num_batches minibatch
updates (loss fwd/bwd + clip + Adam).  Workload = BASELINE.json configs[1]: CartPole,
MLPActorCriticPolicy defaults (12 995 parameters), 65 536 envs x 128 steps per GPU,
10 epochs x 32 minibatches.  With N GPUs every rank owns its own 65 536-env shard (weak
scaling) and the minibatch gradients are all-reduced over NCCL.

The reference (pure Python/JAX) cannot run in this image (no jax wheels, no network), so the
reference arm times the repo's NumPy restatement of the reference (oracle/, fp32) on the
host cores, on a bounded sample of the same workload -- see DESIGN.md "Measurement".
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
N_ENVS, T_STEPS, N_EPOCHS, N_BATCHES = 65536, 128, 10, 32
WORKLOAD = "CartPole PPO, MLPActorCriticPolicy defaults, 65536 envs x 128 steps per GPU, 10 epochs x 32 minibatches"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=N_ENVS, help="envs per GPU")
    ap.add_argument("--horizon", type=int, default=T_STEPS)
    ap.add_argument("--epochs", type=int, default=N_EPOCHS)
    ap.add_argument("--batches", type=int, default=N_BATCHES)
    ap.add_argument("--cpu-envs", type=int, default=4096, help="env count of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------- helpers
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], tflops=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    tflops_burst=p["bf16_tflops"], source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tflops=1400.0, tflops_burst=1590.0, source="fallback (B200_PROFILING.md)")


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
def cpu_iteration(n_envs, T, epochs, batches, seed=0):
    """One PPO iteration of the NumPy restatement of the reference; returns seconds."""
    import numpy as np

    from oracle import envs as OE, gae as OG, policy as OP, ppo as OU, rollout as OR

    oenv = OE.cartpole()
    spec = OP.make_spec(4, 2, False)
    params = OP.init_params(spec, seed=seed)
    rng = np.random.default_rng(seed + 1)
    y = OE.initial_from_uniform(oenv, rng.random((4, n_envs)).astype(np.float32))
    na = rng.gumbel(size=(T, n_envs, 2)).astype(np.float32)
    ru = rng.random((T, n_envs, 4)).astype(np.float32)
    idx = np.stack([OU.batch_indices(n_envs * T, n_envs * T // batches, rng) for _ in range(epochs)])
    t0 = time.perf_counter()
    ro = OR.rollout(oenv, spec, params, y, np.zeros(n_envs, np.float32), np.zeros(n_envs, np.int32), T, 0.99, na, ru)
    adv, ret = OG.gae(ro["rew"], ro["val"], ro["done"], ro["last_value"], 0.99, 0.95)
    flat = {k: OU.flatten_time_major(v) for k, v in
            dict(obs=ro["obs"], act=ro["act"], logp=ro["logp"], val=ro["val"], ret=ret, adv=adv).items()}
    OU.train(spec, params, OU.AdamState.init(params), flat, idx, OU.Hyper())
    return time.perf_counter() - t0


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info

        n = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        n = 1
    return max(n, 1)


def cpu_baseline(args, iters=1):
    import numpy as np  # noqa: F401  (loads BLAS so threadpoolctl sees it)

    n = args.cpu_envs
    cpu_iteration(min(n, 256), args.horizon, 1, args.batches)  # warm caches / BLAS threads
    secs = [cpu_iteration(n, args.horizon, args.epochs, args.batches, seed=i) for i in range(iters)]
    best = min(secs)
    return {
        "value": n * args.horizon / best, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
        "sample": f"{n} envs x {args.horizon} steps, {args.epochs} epochs x {args.batches} minibatches "
                  f"(1/{max(args.envs // n, 1)} of the per-GPU workload), NumPy fp32 oracle, {best:.2f} s/iteration; "
                  f"host has {os.cpu_count()} logical cores; real Lerax/JAX is not installable in this image",
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_envs
    for _ in range(min(args.warmup, 1)):
        cpu_iteration(min(n, 256), args.horizon, 1, args.batches)
    secs = [cpu_iteration(n, args.horizon, args.epochs, args.batches, seed=i) for i in range(args.steps)]
    ms = 1e3 * sum(secs) / len(secs)
    value = n * args.horizon / (ms / 1e3)
    cb = {"value": value, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
          "sample": f"{n} envs x {args.horizon} steps per step (bounded sample of the {args.envs}-env workload), "
                    f"NumPy fp32 restatement of the reference (JAX not installable offline)"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_gpu": args.envs, "horizon": args.horizon, "epochs": args.epochs,
                   "minibatches": args.batches, "sample_envs": n},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from lerax_b200 import _lib as L
    from lerax_b200 import random as jr
    from lerax_b200.algorithm import PPO
    from lerax_b200.env.classic_control import CartPole
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

    env = CartPole()
    n_total = args.envs * world
    algo = PPO(num_envs=n_total, num_steps=args.horizon, num_epochs=args.epochs, num_batches=args.batches)
    policy = MLPActorCriticPolicy(env, key=jr.key(0))
    state = algo.reset(env, policy, key=jr.key(1), callback=algo.consolidate_callbacks(None))
    cb = algo.consolidate_callbacks(None)
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
    for i, k in enumerate(keys[args.warmup:]):
        rollout_key, train_key, _ = jr.split(k, 3)
        ev[i][0].record(stream)
        b = algo._buffers(env, state.policy, state.policy.params.device)
        out = L.LrxRolloutOut(L.ptr(b["obs"]), L.ptr(b["act"]), L.ptr(b["rew"]), L.ptr(b["done"]), L.ptr(b["logp"]),
                              L.ptr(b["val"]), L.ptr(b["last_value"]), None, None, None)
        rng = L.LrxRng(state.rollout_seed, rank * args.envs, state.steps_done & 0xFFFFFFFF)
        L.check(L.lib.lrx_rollout(env.desc(), state.policy.desc(), L.ptr(state.policy.params),
                                  state.step_state.env_state.c_state(), rng, None, out, args.envs, args.horizon,
                                  float(algo.gamma), L.stream_ptr()), "lrx_rollout")
        ev[i][1].record(stream)
        from lerax_b200.buffer import RolloutBuffer

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
        return

    peaks = measured_peaks()
    fwd, dw, dx = policy_flops_per_row(policy._dims)
    rows = args.envs * args.horizon  # per GPU
    upd_flops = 2.0 * (fwd + dw + dx) * rows * args.epochs
    roll_flops = 2.0 * fwd * rows + 300.0 * rows
    row_bytes = 4 * policy.obs_dim + 4 + 4 + 1 + 4 + 4
    stages = {
        "rollout": {"ms": roll_ms, "env_steps_per_s": rows / (roll_ms / 1e3), "tflops": roll_flops / (roll_ms / 1e3) / 1e12,
                    "hbm_gbs": rows * row_bytes / (roll_ms / 1e3) / 1e9, "bytes_per_env_step": row_bytes},
        "gae": {"ms": gae_ms, "hbm_gbs": rows * 17 / (gae_ms / 1e3) / 1e9, "frac_of_hbm_peak": rows * 17 / (gae_ms / 1e3) / 1e9 / peaks["hbm_gbs"],
                "bytes_per_element": 17},
        "update": {"ms": upd_ms, "tflops": upd_flops / (upd_ms / 1e3) / 1e12, "flop_per_row_epoch": 2 * (fwd + dw + dx)},
    }
    dominant = max(("rollout", roll_ms), ("gae", gae_ms), ("update", upd_ms), key=lambda x: x[1])[0]
    if dominant == "gae":
        roof = {"bound": "hbm", "achieved": stages["gae"]["hbm_gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s"}
    else:
        ach = stages[dominant]["tflops"]
        roof = {"bound": "tensor", "achieved": ach, "peak": peaks["tflops"], "unit": "TFLOP/s"}
    notes = {
        "update": "ppo_grad_fused_kernel (tcgen05, bf16x3 fp32 emulation): achieved = algorithmic fp32 FLOPs (75.3 kFLOP "
                  "per row-epoch) / update-stage time. The kernel executes 6 bf16 MMAs per fp32 product and pads 16-wide "
                  "layers to the M = 64 tile: 384 MMAs = 29.1 MFLOP per 64-row tile = 455 kFLOP per row-epoch, i.e. "
                  "`executed_tflops` (ncu: tensor pipe 28 % math-active). traffic = dram bytes per fused launch (one "
                  "minibatch of envs*horizon/minibatches rows; 9.4 MB algorithmic, the rest is one 32-byte sector per "
                  "gathered 4-byte plane element) from the ncu --set full capture in profiles/r1_ppo_grad_fused.md",
        "rollout": "rollout_tc_kernel (tcgen05, bf16x3 fp32 emulation): algorithmic fp32 FLOPs / rollout time against the "
                   "measured bf16 tensor peak",
        "gae": "gae_tn2_kernel: 17 algorithmic bytes per element",
    }
    traffic = {"update": 131.0e6 if (args.envs, args.horizon, args.batches) == (N_ENVS, T_STEPS, N_BATCHES) else None}
    roof.update({"frac": roof["achieved"] / roof["peak"], "traffic": traffic.get(dominant), "kernel": dominant,
                 "peak_source": peaks["source"], "note": notes[dominant]})
    if dominant == "update" and [list(c) for c in policy._dims] == [[4, 64, 64, 16], [16, 64, 64, 1], [16, 64, 16, 2]]:
        executed = 455.0e3 * rows * args.epochs / (upd_ms / 1e3) / 1e12  # bf16 MMA FLOPs actually issued
        roof.update({"executed_tflops": executed, "executed_frac": executed / peaks["tflops"]})

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_gpu": args.envs, "horizon": args.horizon, "epochs": args.epochs,
                   "minibatches": args.batches, "params": policy.n_params,
                   "l2": "rollout buffer (%.0f MB per GPU) exceeds the 126 MB L2; no explicit flush" % (rows * (row_bytes + 8) / 1e6),
                   "parallelism": f"env-sharded x{world}, per-minibatch gradient all-reduce" if world > 1 else "single GPU"},
        "clocks": clocks, "gpu_launches": launches, "stages": stages, "roofline": roof,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
