"""ScheduledCurriculum through PPO.learn: mirrors of tests/test_curriculum.py of the reference, plus the check that
the rewritten field really reaches the kernels (physical parameters are runtime arguments of every launch)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(seed):
    from lerax_b200 import random as jr
    from lerax_b200.env.classic_control import Pendulum
    from lerax_b200.policy import MLPActorCriticPolicy

    policy_key, learn_key = jr.split(jr.key(seed), 2)
    env = Pendulum()
    return env, MLPActorCriticPolicy(env=env, key=policy_key), learn_key


def test_scheduled_curriculum_linear():  # test_curriculum.py:11-29
    from lerax_b200.algorithm import PPO
    from lerax_b200.curriculum import ScheduledCurriculum, linear_schedule

    env, policy, learn_key = _setup(0)
    curriculum = ScheduledCurriculum(where=lambda env: env.m, schedule_fn=linear_schedule(start=0.5, end=2.0, total=10))
    algo = PPO(num_envs=1, num_steps=32)
    trained = algo.learn(env, policy, total_timesteps=320, key=learn_key, callback=curriculum)
    assert trained is not None
    # 10 iterations: the last apply_curriculum saw iteration_count = 10 -> m = 2.0; the caller's env is untouched
    assert algo.last_state.iteration_count == 10
    assert algo.last_state.env.m == pytest.approx(2.0) and env.m == 1.0
    assert algo.last_state.env.desc().p[3] == np.float32(2.0)  # what the next kernel launch would receive


def test_scheduled_curriculum_step():  # test_curriculum.py:32-53
    from lerax_b200.algorithm import PPO
    from lerax_b200.curriculum import ScheduledCurriculum, step_schedule

    env, policy, learn_key = _setup(1)
    curriculum = ScheduledCurriculum(where=lambda env: env.g,
                                     schedule_fn=step_schedule(values=[5.0, 7.0, 9.8], boundaries=[3, 7]))
    algo = PPO(num_envs=1, num_steps=32)
    trained = algo.learn(env, policy, total_timesteps=320, key=learn_key, callback=curriculum)
    assert trained is not None and algo.last_state.env.g == pytest.approx(9.8)


def test_scheduled_curriculum_with_other_callbacks():  # test_curriculum.py:56-72
    from lerax_b200.algorithm import PPO
    from lerax_b200.callback import CallbackList
    from lerax_b200.curriculum import ScheduledCurriculum, linear_schedule

    env, policy, learn_key = _setup(2)
    curriculum = ScheduledCurriculum(where=lambda env: env.m, schedule_fn=linear_schedule(start=0.5, end=2.0, total=10))
    algo = PPO(num_envs=1, num_steps=32)
    trained = algo.learn(env, policy, total_timesteps=320, key=learn_key, callback=CallbackList(callbacks=[curriculum]))
    assert trained is not None and algo.last_state.env.m == pytest.approx(2.0)


def test_curriculum_changes_the_dynamics_the_rollout_sees():
    """Same seeds, gravity rewritten to 0 after the first iteration: the second rollout's rewards differ from the
    run without a curriculum, the first rollout's do not."""
    import torch

    from lerax_b200.algorithm import PPO
    from lerax_b200.callback import AbstractCallback
    from lerax_b200.curriculum import ScheduledCurriculum, step_schedule

    class Tap(AbstractCallback):
        def __init__(self):
            self.rewards = []

        def on_iteration(self, ctx, *, key=None):
            self.rewards.append(ctx.algorithm._bufs["rew"].clone())
            return ctx.state

    runs = []
    for with_curriculum in (False, True):
        env, policy, learn_key = _setup(3)
        tap = Tap()
        cbs = [tap]
        if with_curriculum:
            cbs.append(ScheduledCurriculum(where=lambda env: env.g,
                                           schedule_fn=step_schedule(values=[9.8, 0.0], boundaries=[1])))
        algo = PPO(num_envs=8, num_steps=32, num_epochs=1, num_batches=2)
        algo.learn(env, policy, total_timesteps=8 * 32 * 2, key=learn_key, callback=cbs)
        torch.cuda.synchronize()
        runs.append(tap.rewards)
    assert torch.equal(runs[0][0], runs[1][0])
    assert not torch.equal(runs[0][1], runs[1][1])


def test_adaptive_curriculum_step_state_matches_reference_recurrence():
    """AdaptiveCurriculumStepState (curriculum/adaptive.py:64-70) replayed from the rollout's (reward, done) buffers ==
    the step-by-step accumulation (oracle/callback.py), bit for bit; metric_fn sees the whole [T][N] buffers."""
    import torch

    from lerax_b200.curriculum import AdaptiveCurriculumStepState, LevelCurriculum
    from oracle import callback as OC

    rng = np.random.default_rng(3)
    T, N = 37, 21
    rew = rng.standard_normal((T, N)).astype(np.float32)
    done = rng.random((T, N)) < 0.15
    m0 = rng.standard_normal(N).astype(np.float32)
    cur = LevelCurriculum(where=lambda env: env.max_speed, levels=[4.0, 6.0, 8.0],
                          metric_fn=lambda d, r, loc: r * 2.0 + d.to(r.dtype), threshold=1.0)
    st = AdaptiveCurriculumStepState(torch.from_numpy(m0).cuda())
    out = cur.advance_step_state(st, torch.from_numpy(rew).cuda(), torch.from_numpy(done.astype(np.uint8)).cuda(), N, T)
    want = OC.adaptive_curriculum_step_state(m0, rew * np.float32(2.0) + done.astype(np.float32), done)
    np.testing.assert_array_equal(out.episode_metric.cpu().numpy(), want)


def test_level_curriculum_through_learn():
    """LevelCurriculum (adaptive.py:121-173) in PPO.learn: the env field steps through the levels as the running metric
    crosses the threshold, and the rewritten field reaches the kernels."""
    from lerax_b200.algorithm import PPO
    from lerax_b200.curriculum import LevelCurriculum

    env, policy, learn_key = _setup(4)
    # Pendulum rewards are <= 0: the metric -reward is >= 0 and accumulates to well above the threshold within an iteration
    cur = LevelCurriculum(where=lambda env: env.max_speed, levels=[4.0, 6.0, 8.0], metric_fn=lambda d, r, loc: -r,
                          threshold=5.0, smoothing=0.5)
    algo = PPO(num_envs=2, num_steps=32, num_epochs=1, num_batches=2)
    trained = algo.learn(env, policy, total_timesteps=2 * 32 * 1, key=learn_key, callback=cur)
    assert trained is not None
    st = algo.last_state  # one iteration: the running metric 0.5 * (sum of -reward) crossed 5 -> level 0 -> 1
    assert st.callback_state.level == 1 and st.env.max_speed == pytest.approx(6.0) and env.max_speed == 8.0
    assert st.env.desc().p[0] == np.float32(6.0)  # what the next rollout launch receives
    assert float(st.step_state.callback_state.episode_metric.min()) >= 0.0
    algo.learn(env, policy, total_timesteps=2 * 32 * 5, key=learn_key, callback=cur)
    assert algo.last_state.callback_state.level == 2 and algo.last_state.env.max_speed == pytest.approx(8.0)  # capped at the last level
