"""The callback-state oracle against a scalar transcription of LoggingCallbackStepState.next
(reference src/lerax/callback/logging/callback.py:139-176) -- one env, one step at a time."""
import numpy as np

from oracle import callback as ocb


def _scalar_next(s, reward, done, alpha):
    f = np.float32
    episode_return = f(s["episode_return"] * (f(1.0) - f(s["episode_done"])) + f(reward))
    episode_length = np.int32(s["episode_length"] * (1 - int(s["episode_done"])) + 1)
    if done:
        average_return = f(f(alpha) * episode_return + (f(1.0) - f(alpha)) * s["average_return"])
        average_length = f(f(alpha) * f(episode_length) + (f(1.0) - f(alpha)) * s["average_length"])
    else:
        average_return, average_length = s["average_return"], s["average_length"]
    return {"step": s["step"] + 1, "episode_return": episode_return, "episode_length": episode_length,
            "episode_done": bool(done), "average_return": average_return, "average_length": average_length}


def test_scan_matches_scalar_transcription():
    rng = np.random.default_rng(0)
    T, N, alpha = 97, 13, 0.9
    rew = rng.normal(size=(T, N)).astype(np.float32)
    done = rng.random((T, N)) < 0.08
    out = ocb.scan(ocb.initial(N), rew, done, alpha)
    for n in range(N):
        s = {"step": 0, "episode_return": np.float32(0), "episode_length": np.int32(0), "episode_done": False,
             "average_return": np.float32(0), "average_length": np.float32(0)}
        for t in range(T):
            s = _scalar_next(s, rew[t, n], done[t, n], alpha)
        for k, v in s.items():
            assert out[k][n] == v, (k, n)


def test_known_answer_two_episodes():
    # rewards 1,1,1 (done) then 2,2 (done), alpha = 0.5: EMAs 0.5*3 = 1.5 -> 0.5*4 + 0.5*1.5 = 2.75
    rew = np.array([[1], [1], [1], [2], [2], [5]], np.float32)
    done = np.array([[0], [0], [1], [0], [1], [0]], bool)
    s = ocb.scan(ocb.initial(1), rew, done, 0.5)
    assert s["average_return"][0] == np.float32(2.75)
    assert s["average_length"][0] == np.float32(0.5 * 2 + 0.5 * 1.5)
    assert s["episode_return"][0] == np.float32(5) and s["episode_length"][0] == 1 and s["step"][0] == 6
    assert ocb.iteration_scalars(s) == {"episode/return": 2.75, "episode/length": 1.75, "step": 6}


def test_split_scan_equals_one_scan():
    rng = np.random.default_rng(1)
    rew = rng.normal(size=(40, 7)).astype(np.float32)
    done = rng.random((40, 7)) < 0.2
    a = ocb.scan(ocb.initial(7), rew, done, 0.9)
    b = ocb.scan(ocb.scan(ocb.initial(7), rew[:17], done[:17], 0.9), rew[17:], done[17:], 0.9)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
