"""CUDA path (through the C ABI) == the reference's own source.

Same fixtures as tests/test_oracle_refexec.py (tests/golden/refexec_*.npz: the unmodified reference modules executed
over the stand-ins of oracle/_refexec), compared directly with the kernels: no oracle/ in between.  Tolerances are
north_star's: trajectories, log-probs, values, advantages 1e-5 relative in fp32 at the first step and under short
free-running drift; one PPO update 1e-4."""

import numpy as np
import pytest

import helpers as H
import refexec_util as RU

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 2e-6


@pytest.fixture(scope="module", params=RU.CASE_NAMES)
def case(request):
    return RU.Case(request.param)


def lib_objects(case):
    env = H.lib_env(case.kind, max_episode_steps=case.mes)
    return env, H.lib_policy(env, case.params, **case.pol_kw)


def test_env_step_and_reset(case):
    """lrx_env_step / lrx_env_observe / lrx_env_reset vs env.transition / reward / terminal / observation / initial."""
    import torch

    from lerax_b200 import _lib as L

    fx = case.fx
    env = H.lib_env(case.kind)
    y = fx["probe_y"].T.copy()
    M, k = fx["probe_y"].shape
    yd, td, sd = H.dev(y), H.dev(fx["probe_t"]), torch.zeros(M, dtype=torch.int32, device="cuda")
    st = L.LrxEnvState(L.ptr(yd), L.ptr(td), L.ptr(sd))
    obs = torch.empty((M, fx["probe_obs"].shape[1]), dtype=torch.float32, device="cuda")
    L.check(L.lib.lrx_env_observe(env.desc(), st, L.ptr(obs), None, M, L.stream_ptr()), "lrx_env_observe")
    H.assert_close(obs.cpu().numpy(), fx["probe_obs"], 2e-6, 2e-6, "observation")
    act = H.dev(fx["probe_act"])
    rew = torch.empty(M, dtype=torch.float32, device="cuda")
    term = torch.empty(M, dtype=torch.uint8, device="cuda")
    L.check(L.lib.lrx_env_step(env.desc(), st, L.ptr(act), L.ptr(rew), L.ptr(term), None, None, M, L.stream_ptr()),
            "lrx_env_step")
    y1 = yd.cpu().numpy().T
    tol = 1e-4 if case.kind == "acrobot" else 3e-6  # ill-conditioned at dt = 0.2, see tests/test_oracle_refexec.py
    d = np.abs(y1 - fx["probe_y1"])
    if case.kind in ("pendulum", "acrobot"):
        ang = slice(0, 1) if case.kind == "pendulum" else slice(0, 2)
        d[:, ang] = np.minimum(d[:, ang], 2 * np.pi - d[:, ang])
    assert (d <= tol * (1 + np.abs(fx["probe_y1"]))).all(), d.max()
    H.assert_close(td.cpu().numpy(), fx["probe_t1"], 1e-6, 1e-6, "t + dt")
    # flags and rewards: wherever the kernel's own next state is not within round-off of a threshold
    same = d.max(axis=1) < 1e-6
    assert (term.cpu().numpy().astype(bool)[same] == fx["probe_term"].astype(bool)[same]).all()
    H.assert_close(rew.cpu().numpy()[same], fx["probe_rew"][same], 2e-5, 2e-5, "reward")
    u = H.dev(fx["probe_initial_u"][None, :].copy())
    y0 = torch.empty((k, 1), dtype=torch.float32, device="cuda")
    st0 = L.LrxEnvState(L.ptr(y0), L.ptr(torch.empty(1, dtype=torch.float32, device="cuda")),
                        L.ptr(torch.empty(1, dtype=torch.int32, device="cuda")))
    L.check(L.lib.lrx_env_reset(env.desc(), st0, 1, L.LrxRng(0, 0, 0), L.ptr(u), L.stream_ptr()), "lrx_env_reset")
    H.assert_close(y0.cpu().numpy()[:, 0], fx["probe_initial"], 1e-6, 1e-7, "initial")


def test_rollout_and_gae(case):
    """lrx_rollout + lrx_gae vs PPO.collect_rollout (PPO.step scanned, GAE.__call__) on the same injected noise."""
    fx = case.fx
    env, pol = lib_objects(case)
    T = fx["rew"].shape[0]
    got = H.run_rollout(env, pol, fx["y0"], fx["t0"], fx["sc0"], T, case.gamma, fx["action_noise"], fx["reset_u"], trace=False)
    assert (got["done"] == fx["done"].astype(bool)).all()
    if pol.is_box:
        H.assert_close(got["act"], fx["act"], RTOL, 1e-5, "act")
    else:
        assert (got["act"] == fx["act"]).all()
    for k in ("obs", "logp", "val", "rew"):
        H.assert_close(got[k][0], fx[k][0], RTOL, ATOL, k + "[t=0]")   # no accumulated drift: the 1e-5 bar
        H.assert_close(got[k], fx[k], 1e-4, 1e-4, k)                    # 24 free-running steps of unstable systems
    H.assert_close(got["last_value"], fx["last_value"], 1e-4, 1e-4, "last_value")
    if case.mes:
        assert (got["step_count"] == fx["sc_end"]).all()
    # GAE from the reference's own reward / value / done planes: one pass, 1e-5
    adv, ret = H.run_gae(fx["rew"], fx["val"], fx["done"].astype(np.uint8), fx["last_value"], case.gamma, case.lam)
    H.assert_close(adv, fx["adv"], RTOL, 1e-5, "advantages")
    H.assert_close(ret, fx["ret"], RTOL, 1e-5, "returns")


@pytest.mark.parametrize("h", [0, 1, 2])
def test_ppo_loss_gradient(case, h):
    """lrx_ppo_grad vs PPO.ppo_loss_grad on the reference's own buffer and minibatch indices."""
    fx = case.fx
    env, pol = lib_objects(case)
    hp = case.hyper(h)
    buf = H.DeviceBuffer(**case.planes())
    gb, _ = H.run_grad(pol, buf, fx[f"loss_idx_{h}"], hp)
    P = case.spec.n_params
    gg = gb[:P].copy()
    if pol.is_box:
        gg[-1] -= hp.entropy_loss_coefficient  # added once, in lrx_ppo_apply
    want = fx[f"loss_grad_{h}"]
    H.assert_close(gg, want, 1e-4, 1e-5 * np.abs(want).max(), "gradient")
    assert np.linalg.norm(gg - want) <= 1e-4 * np.linalg.norm(want)
    stats = np.array([gb[P] - 1, 0, -gb[P + 1], gb[P + 2] / 2, -gb[P + 3]])
    stats[1] = stats[2] + stats[3] * hp.value_loss_coefficient + stats[4] * hp.entropy_loss_coefficient
    H.assert_close(stats, fx[f"loss_stats_{h}"], 1e-4, 1e-5, "approx_kl, loss, policy, value, entropy")


@pytest.mark.parametrize("h", [0, 1])
def test_ppo_train(case, h):
    """lrx_ppo_update vs PPO.train (2 epochs x 3 minibatches, optax clip_by_global_norm + adam): north_star's 1e-4."""
    fx = case.fx
    env, pol = lib_objects(case)
    hp = case.hyper(h, learning_rate=1e-3)
    gp, gm, gv, cnt, stats, flag = H.run_update(pol, H.DeviceBuffer(**case.planes()), case.train_idx(), hp)
    assert flag == 0 and cnt == int(fx[f"train_count_{h}"])
    # north_star's 1e-4 on parameters, plus -- for the handful of parameters whose gradient is cancellation noise -- what
    # the gradient bar allows through Adam's normalised step (helpers.adam_conditioned_tolerance)
    tol, tight = H.adam_conditioned_tolerance(case.spec, case.params, case.flat(), case.train_idx(), hp, atol=1e-5)
    H.assert_params_close(gp, fx[f"train_params_{h}"], tol, tight, "parameters after PPO.train")
    if h == 0:
        H.assert_close(gm, fx["train_mu_0"], 1e-3, 1e-5 * np.abs(fx["train_mu_0"]).max(), "adam mu")
        H.assert_close(gv, fx["train_nu_0"], 2e-3, 1e-5 * np.abs(fx["train_nu_0"]).max(), "adam nu")
    s = stats.mean(axis=1).mean(axis=0)
    H.assert_close(s[:5], fx[f"train_log_{h}"][:5], 5e-4, 2e-5, "log dict")
