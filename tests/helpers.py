"""Shared helpers for the GPU parity tests: NumPy <-> device marshalling and thin wrappers
that call the C ABI (lerax_b200._lib) exactly as a reference-side binding would."""

from __future__ import annotations

import numpy as np

from oracle import envs as OE
from oracle import policy as OP


def oracle_env(kind: str, **kw):
    return {"cartpole": OE.cartpole, "pendulum": OE.pendulum, "acrobot": OE.acrobot,
            "mountain_car": OE.mountain_car, "continuous_mountain_car": OE.continuous_mountain_car}[kind](**kw)


def oracle_policy_spec(env_spec, **kw):
    n_out = 1 if env_spec.is_box else env_spec.n_actions
    return OP.make_spec(env_spec.obs_dim, n_out, env_spec.is_box, **kw)


def lib_env(kind: str, **kw):
    from lerax_b200.env.classic_control import Acrobot, CartPole, ContinuousMountainCar, MountainCar, Pendulum
    from lerax_b200.wrapper import TimeLimit

    mes = kw.pop("max_episode_steps", 0)
    env = {"cartpole": CartPole, "pendulum": Pendulum, "acrobot": Acrobot, "mountain_car": MountainCar,
           "continuous_mountain_car": ContinuousMountainCar}[kind](**kw)
    return TimeLimit(env, mes) if mes else env


def lib_policy(env, params_np, **kw):
    from lerax_b200.policy import MLPActorCriticPolicy

    return MLPActorCriticPolicy.from_numpy(env, params_np, **kw)


def dev(x):
    import torch

    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def run_rollout(env, policy, y0, t0, sc0, T, gamma, noise_action=None, reset_uniform=None, seed=0,
                env_offset=0, step_offset=0, trace=True, workspace=True):
    """Call lrx_rollout; returns a dict of NumPy arrays shaped like oracle.rollout's output."""
    import torch

    from lerax_b200 import _lib as L

    k, N = y0.shape
    d = policy.obs_dim
    y, t, sc = dev(y0.astype(np.float32)), dev(t0.astype(np.float32)), dev(sc0.astype(np.int32))
    f32 = dict(dtype=torch.float32, device="cuda")
    obs = torch.empty((T, N, d), **f32)
    act = torch.empty((T, N), dtype=torch.float32 if policy.is_box else torch.int32, device="cuda")
    rew, logp, val = (torch.empty((T, N), **f32) for _ in range(3))
    done = torch.empty((T, N), dtype=torch.uint8, device="cuda")
    last = torch.empty((N,), **f32)
    ytr = torch.empty((T, k, N), **f32) if trace else None
    ttr = torch.empty((T, N), **f32) if trace else None
    yn = torch.empty((T, k, N), **f32) if trace else None
    ws_bytes = int(L.lib.lrx_rollout_workspace_bytes(policy.desc(), N)) if workspace else 0
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device="cuda") if workspace else None
    out = L.LrxRolloutOut(L.ptr(obs), L.ptr(act), L.ptr(rew), L.ptr(done), L.ptr(logp), L.ptr(val), L.ptr(last),
                          L.ptr(ytr), L.ptr(ttr), L.ptr(yn), L.ptr(ws), ws_bytes)
    rp = None
    keep = []
    if noise_action is not None or reset_uniform is not None:
        na = dev(noise_action.astype(np.float32)) if noise_action is not None else None
        ru = dev(reset_uniform.astype(np.float32)) if reset_uniform is not None else None
        keep = [na, ru]
        rp = L.LrxReplay(L.ptr(na), L.ptr(ru))
    st = L.LrxEnvState(L.ptr(y), L.ptr(t), L.ptr(sc))
    L.check(L.lib.lrx_rollout(env.desc(), policy.desc(), L.ptr(policy.params), st, L.LrxRng(seed, env_offset, step_offset),
                              rp, out, N, T, float(gamma), L.stream_ptr()), "lrx_rollout")
    torch.cuda.synchronize()
    del keep
    res = dict(obs=obs, act=act, rew=rew, done=done, logp=logp, val=val, last_value=last, y=y, t=t, step_count=sc)
    if trace:
        res.update(y_trace=ytr, t_trace=ttr, y_next_trace=yn)
    res = {k_: v.cpu().numpy() for k_, v in res.items()}
    res["done"] = res["done"].astype(bool)
    return res


def run_gae(rew, val, done, last, gamma, lam, layout="TN", want_ev=False):
    import torch

    from lerax_b200 import _lib as L

    if layout == "TN":
        T, N = rew.shape
        lay = L.LRX_LAYOUT_TN
    else:
        N, T = rew.shape
        lay = L.LRX_LAYOUT_NT
    r, v, d, lv = dev(rew.astype(np.float32)), dev(val.astype(np.float32)), dev(done.astype(np.uint8)), dev(
        np.asarray(last, np.float32).reshape(N))
    adv, ret = torch.empty_like(r), torch.empty_like(r)
    ev = torch.zeros(L.LRX_GAE_EV_DOUBLES, dtype=torch.float64, device="cuda") if want_ev else None
    L.check(L.lib.lrx_gae(L.ptr(r), L.ptr(v), L.ptr(d), L.ptr(lv), L.ptr(adv), L.ptr(ret), N, T, float(gamma),
                          float(lam), lay, L.ptr(ev), L.stream_ptr()), "lrx_gae")
    torch.cuda.synchronize()
    if want_ev:
        return adv.cpu().numpy(), ret.cpu().numpy(), ev[:4].cpu().numpy()
    return adv.cpu().numpy(), ret.cpu().numpy()


class DeviceBuffer:
    """A time-major rollout buffer on the device built from NumPy planes [T, N]."""

    def __init__(self, obs, act, logp, val, ret, adv, packed=True):
        from lerax_b200 import _lib as L

        self.T, self.N = logp.shape
        self.t = dict(obs=dev(obs.astype(np.float32)), act=dev(act), logp=dev(logp.astype(np.float32)),
                      val=dev(val.astype(np.float32)), ret=dev(ret.astype(np.float32)), adv=dev(adv.astype(np.float32)))
        self.view = L.LrxBatchView(L.ptr(self.t["obs"]), L.ptr(self.t["act"]), L.ptr(self.t["logp"]),
                                   L.ptr(self.t["val"]), L.ptr(self.t["ret"]), L.ptr(self.t["adv"]), self.N, self.T, None)
        if packed and self.t["obs"].shape[-1] <= 8:
            self.pack()

    def pack(self):
        """The row-major copy (lrx_ppo_pack_rows) the production path hands to the update; rebuilt after any edit of the
        planes."""
        import torch

        from lerax_b200 import _lib as L

        self.packed = torch.empty((self.N * self.T, L.LRX_PACKED_FLOATS), dtype=torch.float32, device="cuda")
        self.view.packed = None
        d = int(self.t["obs"].shape[-1])
        L.check(L.lib.lrx_ppo_pack_rows(self.view, d, int(self.t["act"].dtype.is_floating_point), L.ptr(self.packed),
                                        L.stream_ptr()), "lrx_ppo_pack_rows")
        self.view.packed = L.ptr(self.packed)


def make_hyper(hp):
    from lerax_b200 import _lib as L

    return L.LrxPpoHyper(hp.clip_coefficient, hp.value_loss_coefficient, hp.entropy_loss_coefficient,
                         hp.max_grad_norm, hp.learning_rate, hp.b1, hp.b2, hp.eps, int(hp.clip_value_loss),
                         int(hp.normalize_advantages), int(getattr(hp, "surrogate", 0)))


def run_grad(policy, buf: DeviceBuffer, idx_row, hp, B_global=None):
    """lrx_ppo_adv_sums + lrx_ppo_grad for one minibatch; returns gradbuf as NumPy."""
    import torch

    from lerax_b200 import _lib as L

    B = len(idx_row)
    idx = dev(np.asarray(idx_row, np.int32))
    d = policy.desc()
    ws_bytes = int(L.lib.lrx_ppo_workspace_bytes(d, B))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    sums = torch.zeros(2, dtype=torch.float64, device="cuda")
    L.check(L.lib.lrx_ppo_adv_sums(buf.view.adv, buf.N, buf.T, L.ptr(idx), 1, B, L.ptr(sums), L.stream_ptr()),
            "lrx_ppo_adv_sums")
    g = torch.zeros(policy.n_params + L.LRX_PPO_NSTATS, dtype=torch.float32, device="cuda")
    h = make_hyper(hp)
    L.check(L.lib.lrx_ppo_grad(d, L.ptr(policy.params), buf.view, L.ptr(idx), B, B_global or B, L.ptr(sums), h,
                               L.ptr(g), L.ptr(ws), ws_bytes, L.stream_ptr()), "lrx_ppo_grad")
    torch.cuda.synchronize()
    return g.cpu().numpy(), sums.cpu().numpy()


def run_update(policy, buf: DeviceBuffer, idx_epochs, hp, m=None, v=None, count=0):
    """lrx_ppo_update over [E][nb][B] indices.  Returns params, m, v, count, stats[E][nb][8], flag."""
    import torch

    from lerax_b200 import _lib as L

    E, nb, B = idx_epochs.shape
    d = policy.desc()
    params = policy.params.clone()
    m = torch.zeros_like(params) if m is None else dev(m.astype(np.float32))
    v = torch.zeros_like(params) if v is None else dev(v.astype(np.float32))
    cnt = torch.full((), count, dtype=torch.int32, device="cuda")
    ws_bytes = int(L.lib.lrx_ppo_workspace_bytes(d, B))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    stats = torch.zeros((E, nb, L.LRX_PPO_NSTATS), dtype=torch.float32, device="cuda")
    flag = torch.zeros((), dtype=torch.int32, device="cuda")
    h = make_hyper(hp)
    for e in range(E):
        idx = dev(np.ascontiguousarray(idx_epochs[e], np.int32))
        L.check(L.lib.lrx_ppo_update(d, L.ptr(params), L.ptr(m), L.ptr(v), L.ptr(cnt), buf.view, L.ptr(idx), nb, B, h,
                                     L.ptr(stats[e]), L.ptr(flag), L.ptr(ws), ws_bytes, L.stream_ptr()),
                "lrx_ppo_update")
    torch.cuda.synchronize()
    return (params.cpu().numpy(), m.cpu().numpy(), v.cpu().numpy(), int(cnt.item()), stats.cpu().numpy(),
            int(flag.item()))


def assert_close(a, b, rtol, atol, what=""):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    if not (err <= tol).all():
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError(f"{what}: max violation at {i}: got {a[i]!r} want {b[i]!r} "
                             f"(|err|={err[i]:.3e}, tol={tol[i]:.3e}); {int((err > tol).sum())}/{err.size} bad")


def adam_conditioned_tolerance(spec, params, flat, idx_epochs, hp, rtol=1e-4, atol=1e-6, grad_rtol=1e-4, grad_afrac=1e-5):
    """Per-parameter tolerance for parameters after the Adam steps of `idx_epochs`, DERIVED FROM THE GRADIENT BAR.

    optax's update m_hat / (sqrt(v_hat) + eps) is homogeneous of degree 0 in the gradients, so a parameter whose
    gradient entries are cancellation noise (|g_i| ~ 1e-9 while max |g| ~ 0.4 happens for dead-ish units) takes steps
    whose direction no fp32 evaluation pins down: a gradient perturbation d moves one step by ~ lr * d / (sqrt(v_hat) +
    eps), at most 2 lr.  With d_i = grad_rtol |g_i| + grad_afrac max|g| (the bar the gradient tests enforce) this adds
    ~lr * 1e-4 per step to well-conditioned parameters and up to 2 lr per step to the handful of ill-conditioned ones.
    Returns (tol, tight) -- the allowance-free tolerance `tight` lets a test bound HOW MANY parameters needed it."""
    from oracle import ppo as OU

    p = params.copy()
    st = OU.AdamState.init(p)
    extra = np.zeros(p.shape, np.float64)
    for idx in idx_epochs:
        for row in idx:
            _, _, g = OU.ppo_loss_and_grad(spec, p, OU.gather(flat, row), hp)
            gc, _ = OU.clip_by_global_norm(g, hp.max_grad_norm)
            p, st = OU.adam_update(p, gc, st, hp)
            d = grad_rtol * np.abs(gc) + grad_afrac * np.abs(gc).max()
            v_hat = st.v.astype(np.float64) / (1.0 - hp.b2 ** st.count)
            extra += hp.learning_rate * np.minimum(2.0, 2.0 * d / (np.sqrt(v_hat) + hp.eps))
    tight = rtol * np.abs(p) + atol
    return tight + extra, tight


def assert_params_close(got, want, tol, tight, what="", max_loose_frac=2e-3):
    err = np.abs(np.asarray(got, np.float64) - np.asarray(want, np.float64))
    if not (err <= tol).all():
        i = int(np.argmax(err - tol))
        raise AssertionError(f"{what}: parameter {i}: got {got[i]!r} want {want[i]!r} (|err|={err[i]:.3e}, tol={tol[i]:.3e}); "
                             f"{int((err > tol).sum())}/{err.size} bad")
    loose = float((err > tight).mean())
    assert loose <= max_loose_frac, f"{what}: {loose:.2%} of the parameters needed the conditioning allowance"
