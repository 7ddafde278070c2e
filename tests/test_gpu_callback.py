"""lrx_episode_stats and the built-in callbacks against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run_kernel(rew, done, alpha, layout_nt=False, state=None):
    import torch

    from lerax_b200 import _lib as L
    from lerax_b200.callback import LoggingCallbackStepState

    T, N = rew.shape
    dev = torch.device("cuda")
    st = state or LoggingCallbackStepState.initial(N, dev)
    r = torch.from_numpy(np.ascontiguousarray(rew.T if layout_nt else rew)).to(dev)
    d = torch.from_numpy(np.ascontiguousarray((done.T if layout_nt else done).astype(np.uint8))).to(dev)
    L.check(L.lib.lrx_episode_stats(L.ptr(r), L.ptr(d), st.c_struct(), N, T, float(alpha),
                                    L.LRX_LAYOUT_NT if layout_nt else L.LRX_LAYOUT_TN, L.stream_ptr()), "episode_stats")
    torch.cuda.synchronize()
    return st


def _assert_state_equal(st, ref):
    for k, v in ref.items():
        got = getattr(st, k).cpu().numpy()
        assert np.array_equal(got.astype(v.dtype), v), k


@pytest.mark.parametrize("layout_nt", [False, True])
def test_kernel_bit_exact_vs_oracle(layout_nt):
    from oracle import callback as ocb

    rng = np.random.default_rng(3)
    T, N, alpha = 61, 1000, 0.9
    rew = rng.normal(size=(T, N)).astype(np.float32)
    done = rng.random((T, N)) < 0.05
    st = _run_kernel(rew, done, alpha, layout_nt)
    ref = ocb.scan(ocb.initial(N), rew, done, alpha)
    _assert_state_equal(st, ref)
    # the state carries over to the next rollout
    rew2 = rng.normal(size=(T, N)).astype(np.float32)
    done2 = rng.random((T, N)) < 0.05
    st = _run_kernel(rew2, done2, alpha, layout_nt, state=st)
    _assert_state_equal(st, ocb.scan(ref, rew2, done2, alpha))


def test_empty_and_single_step():
    import torch

    from lerax_b200 import _lib as L
    from lerax_b200.callback import LoggingCallbackStepState
    from oracle import callback as ocb

    st = LoggingCallbackStepState.initial(5, torch.device("cuda"))
    assert L.lib.lrx_episode_stats(None, None, st.c_struct(), 5, 0, 0.9, L.LRX_LAYOUT_TN, L.stream_ptr()) == 0
    assert L.lib.lrx_episode_stats(None, None, st.c_struct(), 0, 7, 0.9, L.LRX_LAYOUT_TN, L.stream_ptr()) == 0
    assert L.lib.lrx_episode_stats(None, None, st.c_struct(), 5, 7, 0.9, L.LRX_LAYOUT_TN, L.stream_ptr()) != 0
    ref = ocb.initial(5)
    rng = np.random.default_rng(4)
    for _ in range(6):  # LoggingCallbackStepState.next == one on_step
        r = rng.normal(size=5).astype(np.float32)
        d = rng.random(5) < 0.4
        st = st.next(torch.from_numpy(r).cuda(), torch.from_numpy(d).cuda(), 0.7)
        ref = ocb.next_state(ref, r, d, 0.7)
    torch.cuda.synchronize()
    _assert_state_equal(st, ref)


def test_learn_with_logging_and_progress_callbacks():
    """examples/ppo.py with its callbacks: one record per iteration, step = env steps so far, the episode
    statistics equal the oracle's scan over the same reward / done streams."""
    import torch

    from lerax_b200 import random as jr
    from lerax_b200.algorithm import PPO
    from lerax_b200.callback import HistoryBackend, LoggingCallback, ProgressBarCallback
    from lerax_b200.env.classic_control import CartPole
    from lerax_b200.policy import MLPActorCriticPolicy
    from oracle import callback as ocb

    env = CartPole()
    policy = MLPActorCriticPolicy(env, key=jr.key(0))
    N, T, iters = 64, 32, 5
    algo = PPO(num_envs=N, num_steps=T, num_epochs=2, num_batches=4)
    hist = HistoryBackend()
    streams = []

    class Tap(LoggingCallback):  # records the streams the kernel saw
        def advance_step_state(self, step_state, rewards, dones, num_envs, num_steps):
            streams.append((rewards.cpu().numpy().copy(), dones.cpu().numpy().astype(bool).copy()))
            return super().advance_step_state(step_state, rewards, dones, num_envs, num_steps)

    bar = ProgressBarCallback(N * T * iters, env=env, policy=policy, disable=True)
    algo.learn(env, policy, total_timesteps=N * T * iters, key=jr.key(1),
               callback=[Tap(hist, env=env, policy=policy, alpha=0.9), bar])
    torch.cuda.synchronize()
    assert len(hist.records) == iters and bar.completed == N * T * iters
    ref = ocb.initial(N)
    for it, (rew, done) in enumerate(streams):
        ref = ocb.scan(ref, rew, done, 0.9)
        step, scalars = hist.records[it]
        want = ocb.iteration_scalars(ref)
        assert step == want["step"] == N * T * (it + 1)
        assert scalars["episode/return"] == pytest.approx(want["episode/return"], rel=1e-6)
        assert scalars["episode/length"] == pytest.approx(want["episode/length"], rel=1e-6)
        assert {"train/policy_loss", "train/value_loss", "train/learning_rate"} <= set(scalars)
    assert hist.records[-1][1]["episode/length"] > 0
    assert "algorithm.num_envs" in hist.hparams


def test_custom_on_step_still_rejected():
    from lerax_b200.algorithm import PPO
    from lerax_b200.callback import AbstractCallback

    class Mine(AbstractCallback):
        def on_step(self, ctx, *, key=None):
            return ctx.state

    with pytest.raises(NotImplementedError):
        PPO(num_envs=4, num_steps=8).consolidate_callbacks([Mine()])


def test_ppo_with_callbacks_reference_integration(capsys, tmp_path):
    """tests/integration/test_ppo_callbacks.py of the reference: PPO() defaults, one LoggingCallback (not a list)
    with several backends; 512 timesteps are fewer than one iteration of the default 4 x 2048 rollout."""
    from lerax_b200 import random as jr
    from lerax_b200.algorithm import PPO
    from lerax_b200.callback import ConsoleBackend, HistoryBackend, LoggingCallback, TensorBoardBackend
    from lerax_b200.env.classic_control import CartPole
    from lerax_b200.policy import MLPActorCriticPolicy

    policy_key, learn_key = jr.split(jr.key(0), 2)
    env = CartPole()
    policy = MLPActorCriticPolicy(env=env, key=policy_key)
    algo = PPO()
    hist = HistoryBackend()
    logger = LoggingCallback([TensorBoardBackend(log_dir=tmp_path), hist, ConsoleBackend()], env=env, policy=policy)
    trained = algo.learn(env, policy, total_timesteps=512, key=learn_key, callback=logger)
    assert trained is not None and hist.records == [] and hist.name.startswith("MLPActorCriticPolicy_CartPole_")
    # one full iteration of the defaults logs one record to every backend
    trained = algo.learn(env, policy, total_timesteps=4 * 2048, key=learn_key, callback=logger)
    logger.close()
    assert len(hist.records) == 1 and hist.records[0][0] == 4 * 2048
    runs = list(tmp_path.iterdir())  # log_dir / name / events.out.tfevents.*
    assert len(runs) == 1 and runs[0].name == hist.name and len(list(runs[0].iterdir())) > 0
    assert "episode/return=" in capsys.readouterr().out
