"""Worker of tests/test_gpu_p2p.py (one process per GPU, launched with torch.distributed.run):
the peer-memory gradient exchange against the NCCL all-reduce, then a short PPO run with both."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", local)
    from lerax_b200 import _lib as L
    from lerax_b200._p2p import make_exchange

    n = 13003
    ex = make_exchange(n, dist, dev)
    assert ex is not None, "peer access unavailable"
    out = torch.empty(n, dtype=torch.float32, device=dev)
    for it in range(50):  # many exchanges back to back: slot reuse, flag ordering
        g = torch.Generator(device=dev).manual_seed(1000 * it + rank)
        x = torch.randn(n, generator=g, device=dev)
        slot = ex.next_slot()
        # what lrx_ppo_grad does: write the local partial into the own slot
        L.check(L.debug_lib().lrx_debug_copy_f32(slot, L.ptr(x), n, L.stream_ptr()), "copy")
        ex.allreduce(out)
        ref = x.clone()
        dist.all_reduce(ref)
        torch.cuda.synchronize()
        assert torch.allclose(out, ref, rtol=1e-6, atol=1e-6), (it, (out - ref).abs().max().item())
        # bit-identical on every rank (same rank-ordered sum everywhere)
        gathered = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(gathered, out)
        assert all(torch.equal(gathered[0], t) for t in gathered), it
    ex.check()

    # a short PPO run: peer-memory exchange vs NCCL must give the same parameters (up to summation order)
    from lerax_b200 import random as jr
    from lerax_b200.algorithm import PPO
    from lerax_b200.env.classic_control import CartPole
    from lerax_b200.policy import MLPActorCriticPolicy

    def run(disable):
        os.environ["LRX_DISABLE_P2P"] = "1" if disable else "0"
        env = CartPole()
        policy = MLPActorCriticPolicy(env, key=jr.key(0))
        algo = PPO(num_envs=256 * world, num_steps=32, num_epochs=2, num_batches=4)
        trained = algo.learn(env, policy, total_timesteps=256 * world * 32 * 3, key=jr.key(1))
        torch.cuda.synchronize()
        used = getattr(algo, "_ex", None) is not None
        return trained.params.clone(), used

    p_p2p, used = run(False)
    assert used
    p_nccl, used = run(True)
    assert not used
    assert torch.allclose(p_p2p, p_nccl, rtol=1e-4, atol=1e-5), (p_p2p - p_nccl).abs().max().item()
    params = [torch.empty_like(p_p2p) for _ in range(world)]
    dist.all_gather(params, p_p2p)
    assert all(torch.equal(params[0], t) for t in params), "ranks diverged"
    if rank == 0:
        print("P2P_OK", world, float((p_p2p - p_nccl).abs().max()))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
