"""NVLink peer-memory gradient exchange (lrx_p2p_*): needs >= 2 GPUs on the box, skipped otherwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_p2p_exchange_matches_nccl_on_two_gpus():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(here, "p2p_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "P2P_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("mode", ["2", "1", "0", "3", "push"])
@pytest.mark.parametrize("kind", ["cartpole", "pendulum"])
def test_data_parallel_epoch_with_a_one_rank_exchange(kind, mode, monkeypatch):
    """The data-parallel update path on ONE GPU: a one-rank exchange (the rank's own buffer is its only peer) runs the
    same kernels as a multi-GPU job -- lrx_ppo_update_p2p with the fused reduce / exchange / clip + Adam kernel (one flag
    per rank, the default), its per-block-flag and push variants and the three-call structure -- and must reproduce
    lrx_ppo_update on the same minibatches: the sum over one rank is the rank's own gradient."""
    import ctypes as C

    import numpy as np
    import torch

    import helpers as H
    from lerax_b200 import _lib as L
    from lerax_b200._p2p import GradientExchange
    from oracle import ppo as OU
    from test_gpu_parity import synthetic_buffer

    if mode == "push":
        monkeypatch.setenv("LRX_P2P_FUSED", "0")
        monkeypatch.setenv("LRX_P2P_PUSH", "1")
    else:
        monkeypatch.setenv("LRX_P2P_FUSED", mode)
    T, N, nb = 32, 40, 4
    B = T * N // nb
    spec, params, pol, rng, flat, buf = synthetic_buffer(kind, T, N, seed=9)
    hp = OU.Hyper(entropy_loss_coefficient=0.01)
    idx_np = OU.batch_indices(T * N, B, rng)[None]
    want_p, want_m, want_v, want_cnt, want_stats, flag = H.run_update(pol, buf, idx_np, hp)
    assert flag == 0 and want_cnt == nb

    n = pol.n_params + L.LRX_PPO_NSTATS
    own = C.c_void_p()
    handle = (C.c_ubyte * L.LRX_P2P_HANDLE_BYTES)()
    L.check(L.lib.lrx_p2p_alloc(n, C.byref(own), handle), "lrx_p2p_alloc")
    try:
        ex = GradientExchange(n, 0, 1, own.value, [], torch.tensor([own.value], dtype=torch.int64, device="cuda"),
                              torch.zeros((), dtype=torch.int32, device="cuda"))
        d = pol.desc()
        p = pol.with_params(pol.params.clone())
        m, v = torch.zeros_like(p.params), torch.zeros_like(p.params)
        cnt = torch.zeros((), dtype=torch.int32, device="cuda")

        class Opt:
            mu, nu, count = m, v, cnt

        idx = H.dev(np.ascontiguousarray(idx_np[0], np.int32))
        sums = torch.zeros((nb, 2), dtype=torch.float64, device="cuda")
        L.check(L.lib.lrx_ppo_adv_sums(buf.view.adv, buf.N, buf.T, L.ptr(idx), nb, B, L.ptr(sums), L.stream_ptr()), "adv_sums")
        ws = torch.empty(int(L.lib.lrx_ppo_workspace_bytes(d, B)), dtype=torch.uint8, device="cuda")
        stats = torch.zeros((nb, L.LRX_PPO_NSTATS), dtype=torch.float32, device="cuda")
        nonfinite = torch.zeros((), dtype=torch.int32, device="cuda")
        ex.update_epoch(d, p, Opt, buf.view, idx, B, B, sums, H.make_hyper(hp), None, stats, nonfinite, ws)
        torch.cuda.synchronize()
        ex.check()
        assert int(cnt.item()) == nb and int(nonfinite.item()) == 0 and ex.seq == nb
        # same gradients; the global norm is assembled from differently grouped partial sums in the three-call kernels
        H.assert_close(p.params.cpu().numpy(), want_p, 1e-6, 1e-7, "params")
        H.assert_close(m.cpu().numpy(), want_m, 1e-5, 1e-9, "adam m")
        H.assert_close(stats.cpu().numpy()[:, :6], want_stats[0][:, :6], 1e-5, 1e-6, "stats")
    finally:
        L.lib.lrx_p2p_free(own)
