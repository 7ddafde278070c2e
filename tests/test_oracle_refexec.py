"""oracle/ == the reference's own source.

tests/golden/refexec_*.npz hold outputs of the UNMODIFIED reference modules (env dynamics / transition / reward /
terminal / observation / initial, TimeLimit, PPO.step + collect_rollout under filter_vmap, GAE.__call__, ppo_loss +
its gradient, PPO.train with the optax chain) executed over the NumPy/torch stand-ins for jax / equinox / diffrax /
distreqx / optax of oracle/_refexec (tests/golden/make_refexec.py).  Here the NumPy restatement in oracle/ is held to
them at fp32 round-off tolerances: the two evaluations differ only in libm and summation order."""

import numpy as np
import pytest

import refexec_util as RU
from helpers import assert_close
from oracle import envs as OE
from oracle import gae as OG
from oracle import ppo as OU
from oracle import rollout as OR


@pytest.fixture(scope="module", params=RU.CASE_NAMES)
def case(request):
    return RU.Case(request.param)


def test_fixture_files_present():
    assert len(RU.all_fixture_files()) == len(RU.CASE_NAMES) + 1


def test_env_single_transitions(case):
    """env.dynamics / transition (one Tsit5 step + clip) / reward / terminal / observation on scattered states."""
    fx, env = case.fx, case.env_bare
    y = fx["probe_y"].T.copy()
    act = fx["probe_act"]
    assert_close(OE.dynamics(env, y, act).T, fx["probe_dyn"], 2e-6, 2e-6, "dynamics")
    y1, t1 = OE.transition(env, y, fx["probe_t"], act)
    # acrobot integrates with dt = 0.2 through angular velocities of O(10): ill-conditioned, both fp32 evaluations sit
    # ~5e-5 from the float64 twin there
    tol = 1e-4 if case.kind == "acrobot" else 3e-6
    if case.kind in ("pendulum", "acrobot"):  # wrapped angles: compare on the circle
        d = np.abs(y1.T - fx["probe_y1"])
        ang = slice(0, 1) if case.kind == "pendulum" else slice(0, 2)
        d[:, ang] = np.minimum(d[:, ang], 2 * np.pi - d[:, ang])
        assert (d <= tol * (1 + np.abs(fx["probe_y1"]))).all(), d.max()
    else:
        assert_close(y1.T, fx["probe_y1"], tol, tol, "transition")
    assert_close(t1, fx["probe_t1"], 1e-6, 1e-6, "t + dt")
    got_clip = OE.clip_state(env, fx["probe_y_wide"].T.copy()).T
    dc = np.abs(got_clip - fx["probe_clip"])
    if case.kind in ("pendulum", "acrobot"):
        ang = slice(0, 1) if case.kind == "pendulum" else slice(0, 2)
        dc[:, ang] = np.minimum(dc[:, ang], 2 * np.pi - dc[:, ang])
    assert (dc <= 4e-6 * (1 + np.abs(fx["probe_y_wide"]))).all(), dc.max()
    want_term = fx["probe_term"].astype(bool)
    got_term = OE.terminal(env, fx["probe_y1"].T.copy())
    assert (got_term == want_term).all()
    assert_close(OE.reward(env, y, act, fx["probe_y1"].T.copy()), fx["probe_rew"], 3e-6, 3e-6, "reward")
    assert_close(OE.observation(env, y).T, fx["probe_obs"], 2e-6, 2e-6, "observation")
    got0 = OE.initial_from_uniform(env, fx["probe_initial_u"][:, None])[:, 0]
    assert_close(got0, fx["probe_initial"], 1e-6, 1e-7, "initial")


def test_reset_states(case):
    fx = case.fx
    y0 = OE.initial_from_uniform(case.env, fx["init_u"].T)
    assert_close(y0, fx["y0"], 1e-6, 1e-7, "PPO.reset -> env.initial")
    assert (fx["t0"] == 0).all() and (fx["sc0"] == 0).all()


def test_rollout(case):
    """PPO.collect_rollout for 6 envs x 24 steps with the same injected noise: every buffer plane, the carried state and
    last_value."""
    fx = case.fx
    T = fx["rew"].shape[0]
    got = OR.rollout(case.env, case.spec, case.params, fx["y0"], fx["t0"], fx["sc0"], T, case.gamma, fx["action_noise"],
                     fx["reset_u"])
    assert (got["done"] == fx["done"].astype(bool)).all()
    if case.env.is_box:
        assert_close(got["act"], fx["act"], 2e-5, 2e-5, "act")
    else:
        assert (got["act"] == fx["act"]).all()
    for k in ("obs", "rew", "logp", "val"):
        assert_close(got[k], fx[k], 5e-5, 5e-5, k)
    assert_close(got["last_value"], fx["last_value"], 5e-5, 5e-5, "last_value")
    assert_close(got["y"], fx["y_end"], 5e-5, 5e-5, "carried y")
    if case.mes:  # TimeLimitState.step_count; bare envs carry none
        assert (got["step_count"] == fx["sc_end"]).all()
    assert int(fx["done"].sum()) > 0 or case.kind in ("acrobot", "continuous_mountain_car", "pendulum")


def test_gae_on_rollout(case):
    fx = case.fx
    adv, ret = OG.gae(fx["rew"], fx["val"], fx["done"].astype(bool), fx["last_value"], case.gamma, case.lam)
    assert_close(adv, fx["adv"], 2e-5, 2e-5, "advantages")
    assert_close(ret, fx["ret"], 2e-5, 2e-5, "returns")


def test_gae_call_random_strips():
    fx = np.load(RU.all_fixture_files()[RU.all_fixture_files().index(
        [f for f in RU.all_fixture_files() if f.endswith("refexec_gae.npz")][0])])
    for i in range(int(fx["n"])):
        g, lam = [float(x) for x in fx[f"gae{i}_gl"]]
        adv, ret = OG.gae(fx[f"gae{i}_rew"][:, None], fx[f"gae{i}_val"][:, None], fx[f"gae{i}_done"][:, None],
                          fx[f"gae{i}_last"].reshape(1), g, lam)
        assert_close(adv[:, 0], fx[f"gae{i}_adv"], 2e-5, 2e-6, f"gae{i} adv")
        assert_close(ret[:, 0], fx[f"gae{i}_ret"], 2e-5, 2e-6, f"gae{i} ret")


@pytest.mark.parametrize("h", [0, 1, 2])
def test_ppo_loss_and_gradient(case, h):
    """PPO.ppo_loss_grad (eqx.filter_value_and_grad over the reference's ppo_loss) vs the hand-derived gradient."""
    fx = case.fx
    flat = case.flat()
    mb = OU.gather(flat, fx[f"loss_idx_{h}"])
    _, stats, grads = OU.ppo_loss_and_grad(case.spec, case.params, mb, case.hyper(h))
    assert_close(stats, fx[f"loss_stats_{h}"], 2e-5, 2e-6, "approx_kl, loss, policy, value, entropy")
    want = fx[f"loss_grad_{h}"]
    assert_close(grads, want, 1e-4, 2e-6 * np.abs(want).max(), "gradient")
    assert np.linalg.norm(grads - want) <= 2e-5 * np.linalg.norm(want)


@pytest.mark.parametrize("h", [0, 1])
def test_ppo_train(case, h):
    """PPO.train: 2 epochs x 3 minibatches of train_batch (clip_by_global_norm + adam), log dict incl. explained_variance."""
    fx = case.fx
    hp = case.hyper(h, learning_rate=1e-3)
    p, st, log = OU.train(case.spec, case.params.copy(), OU.AdamState.init(case.params), case.flat(), case.train_idx(), hp)
    assert st.count == int(fx[f"train_count_{h}"]) == 6
    assert_close(p, fx[f"train_params_{h}"], 1e-4, 1e-5, "parameters after PPO.train")
    if h == 0:
        assert_close(st.m, fx["train_mu_0"], 1e-3, 1e-6 * np.abs(fx["train_mu_0"]).max(), "adam mu")
        assert_close(st.v, fx["train_nu_0"], 1e-3, 1e-6 * np.abs(fx["train_nu_0"]).max(), "adam nu")
    got = np.array([log[k] for k in ("approx_kl", "loss", "policy_loss", "value_loss", "entropy_loss", "explained_variance")])
    assert_close(got, fx[f"train_log_{h}"], 5e-4, 2e-5, "log dict")
