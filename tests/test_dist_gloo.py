"""World-size-2 checks of the multi-GPU protocol of PPO.train on CPU over gloo.

The data path itself is CUDA-only; what is rank-dependent is HOST logic and the algebra of the
three collectives: (1) one SUM all-reduce per epoch of the per-minibatch advantage sums, so every
rank normalises with the statistics of the WHOLE minibatch (ppo.py:424-427); (2) one SUM
all-reduce per minibatch of [gradient of the mean loss restricted to the rank's rows || stat
sums], each rank scaling by 1/B_global; (3) env sharding N/world with disjoint Philox env
offsets.  Here the oracle stands in for the kernels and the test asserts that the all-reduced
result equals the single-process gradient of the same minibatch."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import envs as OE
from oracle import gae as OG
from oracle import policy as OP
from oracle import ppo as OU
from oracle import rollout as OR


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _minibatch(seed=0, N=16, T=12):
    oenv = OE.cartpole()
    spec = OP.make_spec(4, 2, False)
    params = OP.init_params(spec, seed=seed)
    rng = np.random.default_rng(seed + 1)
    y0 = OE.initial_from_uniform(oenv, rng.random((4, N)).astype(np.float32))
    na = rng.gumbel(size=(T, N, 2)).astype(np.float32)
    ru = rng.random((T, N, 4)).astype(np.float32)
    ro = OR.rollout(oenv, spec, params, y0, np.zeros(N, np.float32), np.zeros(N, np.int32), T, 0.99, na, ru)
    adv, ret = OG.gae(ro["rew"], ro["val"], ro["done"], ro["last_value"], 0.99, 0.95)
    logp_old = (ro["logp"] + rng.normal(0, 0.2, ro["logp"].shape)).astype(np.float32)
    planes = dict(obs=ro["obs"], act=ro["act"], logp=logp_old, val=ro["val"], ret=ret, adv=adv)
    flat = {k: OU.flatten_time_major(v) for k, v in planes.items()}
    return spec, params, flat


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lerax_b200.algorithm import PPO, _dist

        r, w, d = _dist()
        assert (r, w) == (rank, world) and d is not None
        algo = PPO(num_envs=16, num_steps=12, num_epochs=1, num_batches=4)
        assert algo._local_envs() == 8 and algo.batch_size == 48
        with pytest.raises(ValueError):
            PPO(num_envs=7, num_steps=4, num_batches=1)._local_envs()

        spec, params, flat = _minibatch()
        hp = OU.Hyper(entropy_loss_coefficient=0.01)
        B_global = 96
        idx_global = np.random.default_rng(7).permutation(flat["adv"].shape[0])[:B_global]
        shard = idx_global[rank::world]  # any disjoint split of the minibatch rows
        B_local = len(shard)
        # (1) advantage sums: all-reduced once, every rank normalises with the global statistics
        a = flat["adv"][shard].astype(np.float64)
        sums = torch.tensor([a.sum(), (a * a).sum()], dtype=torch.float64)
        dist.all_reduce(sums)
        mean = sums[0].item() / B_global
        std = np.sqrt(max(sums[1].item() / B_global - mean * mean, 0.0))
        mb = OU.gather(flat, shard)
        mb["adv"] = ((mb["adv"] - np.float32(mean)) / (np.float32(std) + np.float32(1.1920929e-07))).astype(np.float32)
        # (2) the rank's share of the gradient of the MEAN loss over B_global rows
        hp_local = OU.Hyper(entropy_loss_coefficient=hp.entropy_loss_coefficient, normalize_advantages=False)
        _, _, g_local = OU.ppo_loss_and_grad(spec, params, mb, hp_local)
        share = torch.from_numpy((g_local.astype(np.float64) * (B_local / B_global)))
        dist.all_reduce(share)
        if rank == 0:
            _, _, g_full = OU.ppo_loss_and_grad(spec, params, OU.gather(flat, idx_global), hp)
            out.put((share.numpy(), g_full.astype(np.float64)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_single_process():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got, want = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    scale = np.abs(want).max()
    np.testing.assert_allclose(got, want, rtol=2e-4, atol=2e-5 * scale)
