"""Host side of the NVLink peer-memory gradient exchange (include/lerax_b200.h, lrx_p2p_*).

One `GradientExchange` per process: allocates the own exchange buffer, swaps the CUDA IPC handles
with the peers through `torch.distributed.all_gather_object` (any host channel would do), opens
the peers' buffers and keeps the device array of the G buffer addresses the kernel indexes.
`LRX_DISABLE_P2P=1` (or a failed IPC set-up, e.g. GPUs without peer access) keeps the NCCL all-reduce;
every rank executes the same collectives during set-up and all ranks take the same decision."""

from __future__ import annotations

import ctypes as C
import os

from . import _lib as L


class GradientExchange:
    def __init__(self, n_floats, rank, world, own, peers, addresses, error):
        self.n, self.rank, self.world = int(n_floats), rank, world
        self._own, self._peers, self.addresses, self.error = own, peers, addresses, error
        self.seq = 0

    def _seq32(self) -> int:
        # 1 .. 2^32 - 2, never 0 (0 = nothing published yet), with an EVEN period: consecutive exchanges always alternate
        # between the two slots (seq & 1), also across the wrap
        return (self.seq - 1) % 0xFFFFFFFE + 1

    def next_slot(self) -> int:
        """Device address the next lrx_ppo_grad writes its gradbuf to."""
        self.seq += 1
        return int(L.lib.lrx_p2p_slot(self._own, self.n, self._seq32()))

    def grad(self, pol_desc, policy, view, idx_row, B_local, B_global, adv_sums, hyper, workspace) -> None:
        """lrx_ppo_grad_p2p: the next minibatch's gradient into this rank's slot, published by the reduction's last block."""
        self.seq += 1
        L.check(L.lib.lrx_ppo_grad_p2p(pol_desc, L.ptr(policy.params), view, L.ptr(idx_row), B_local, B_global,
                                       L.ptr(adv_sums), hyper, self._own, self._seq32(), L.ptr(workspace),
                                       workspace.numel(), L.stream_ptr()), "lrx_ppo_grad_p2p")

    def allreduce(self, out) -> None:
        """Sum of the slots published for the current seq, in rank order, into `out` (float32 [n])."""
        L.check(L.lib.lrx_p2p_allreduce(L.ptr(self.addresses), self.rank, self.world, self.n, L.ptr(out), self._seq32(),
                                        L.ptr(self.error), L.stream_ptr()), "lrx_p2p_allreduce")

    def apply(self, pol_desc, policy, opt_state, hyper, stats_out, nonfinite) -> None:
        """lrx_ppo_apply_p2p for the current seq: peer sum + global-norm clip + Adam in one cooperative launch."""
        torch = L.require_cuda()
        if getattr(self, "_scratch", None) is None:
            nbytes = int(L.lib.lrx_ppo_apply_p2p_scratch_bytes(pol_desc))
            self._scratch = torch.empty((nbytes,), dtype=torch.uint8, device=self.error.device)
        L.check(L.lib.lrx_ppo_apply_p2p(pol_desc, L.ptr(policy.params), L.ptr(opt_state.mu), L.ptr(opt_state.nu),
                                        L.ptr(opt_state.count), L.ptr(self.addresses), self.rank, self.world,
                                        self._seq32(), hyper, L.ptr(stats_out), L.ptr(nonfinite), L.ptr(self.error),
                                        L.ptr(self._scratch), self._scratch.numel(), L.stream_ptr()),
                "lrx_ppo_apply_p2p")

    def grad_apply(self, pol_desc, policy, opt_state, view, idx_row, B_local, B_global, adv_sums, hyper, stats_out,
                   nonfinite, workspace) -> bool:
        """One minibatch in two launches (lrx_ppo_grad_apply_p2p): gradient kernel, then reduce + block-wise exchange +
        clip + Adam.  Returns False (nothing enqueued, seq untouched) for policies the fused gradient kernels do not cover."""
        self.seq += 1
        handled = C.c_int32(0)
        L.check(L.lib.lrx_ppo_grad_apply_p2p(pol_desc, L.ptr(policy.params), L.ptr(opt_state.mu), L.ptr(opt_state.nu),
                                             L.ptr(opt_state.count), view, L.ptr(idx_row), B_local, B_global,
                                             L.ptr(adv_sums), hyper, L.ptr(self.addresses), self._own, self.rank,
                                             self.world, self._seq32(), L.ptr(stats_out), L.ptr(nonfinite),
                                             L.ptr(self.error), L.ptr(workspace), workspace.numel(), L.stream_ptr(),
                                             C.byref(handled)), "lrx_ppo_grad_apply_p2p")
        if not handled.value:
            self.seq -= 1
            return False
        return True

    def update_epoch(self, pol_desc, policy, opt_state, view, idx, B_local, B_global, adv_sums, hyper, lrs, stats_out,
                     nonfinite, workspace) -> None:
        """lrx_ppo_update_p2p: every minibatch of an epoch (gradient -> NVLink exchange -> clip + Adam) in one C call."""
        torch = L.require_cuda()
        if getattr(self, "_scratch", None) is None:
            nbytes = int(L.lib.lrx_ppo_apply_p2p_scratch_bytes(pol_desc))
            self._scratch = torch.empty((nbytes,), dtype=torch.uint8, device=self.error.device)
        nb = int(idx.shape[0])
        L.check(L.lib.lrx_ppo_update_p2p(pol_desc, L.ptr(policy.params), L.ptr(opt_state.mu), L.ptr(opt_state.nu),
                                         L.ptr(opt_state.count), view, L.ptr(idx), nb, B_local, B_global, L.ptr(adv_sums),
                                         hyper, lrs, L.ptr(self.addresses), self._own, self.rank, self.world, self.seq,
                                         L.ptr(stats_out), L.ptr(nonfinite), L.ptr(self.error), L.ptr(self._scratch),
                                         self._scratch.numel(), L.ptr(workspace), workspace.numel(), L.stream_ptr()),
                "lrx_ppo_update_p2p")
        self.seq += nb

    def check(self) -> None:
        if int(self.error.item()) != 0:
            raise RuntimeError("lrx_p2p_allreduce: a peer rank did not publish its gradient within 10 s")

    def close(self) -> None:
        for p in self._peers:
            L.lib.lrx_p2p_close(p)
        self._peers = []
        if self._own:
            L.lib.lrx_p2p_free(self._own)
            self._own = None


def make_exchange(n_floats: int, dist, device):
    """A GradientExchange, or None when peer access is disabled or unavailable (-> NCCL all-reduce)."""
    torch = L.require_cuda()
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = os.environ.get("LRX_DISABLE_P2P", "0") != "1"
    own, handle = None, bytes(L.LRX_P2P_HANDLE_BYTES)
    if ok:
        p = C.c_void_p()
        h = (C.c_ubyte * L.LRX_P2P_HANDLE_BYTES)()
        if L.lib.lrx_p2p_alloc(int(n_floats), C.byref(p), h) == 0:
            own, handle = p.value, bytes(h)
        else:
            ok = False
    gathered = [None] * world
    dist.all_gather_object(gathered, (rank, ok, handle))
    ok = ok and all(g[1] for g in gathered)
    peers, addrs = [], [0] * world
    if ok:
        for r, _ok, hb in gathered:
            if r == rank:
                addrs[r] = own
                continue
            p = C.c_void_p()
            buf = (C.c_ubyte * L.LRX_P2P_HANDLE_BYTES).from_buffer_copy(hb)
            if L.lib.lrx_p2p_open(buf, C.byref(p)) != 0:  # no peer access between these two GPUs
                ok = False
                break
            peers.append(p.value)
            addrs[r] = p.value
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # also orders "every rank has mapped every buffer" before use
    if int(flag.item()) == 0:
        for p in peers:
            L.lib.lrx_p2p_close(p)
        if own:
            L.lib.lrx_p2p_free(own)
        return None
    return GradientExchange(n_floats, rank, world, own, peers,
                            torch.tensor(addrs, dtype=torch.int64, device=device),
                            torch.zeros((), dtype=torch.int32, device=device))
