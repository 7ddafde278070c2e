"""CartPole learning curve of the NumPy oracle's PPO (the restated reference) with the hyper-parameters of
tests/test_gpu_learning.py: 64 envs x 128 steps, 10 epochs x 32 minibatches, 40 iterations.  CPU only."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import envs as OE, gae as OG, policy as OP, ppo as OU, rollout as OR

def evaluate(oenv, spec, params, episodes=64, max_steps=500, seed=0):
    rng = np.random.default_rng(seed)
    y = OE.initial_from_uniform(oenv, rng.random((4, episodes)).astype(np.float32))
    t = np.zeros(episodes, np.float32); alive = np.ones(episodes, bool); total = np.zeros(episodes)
    for _ in range(max_steps):
        _, head = OP.forward(spec, params, y.T.copy())
        act = head.argmax(-1).astype(np.int32)
        y, t = OE.transition(oenv, y, t, act)
        total += alive
        alive &= ~OE.terminal(oenv, y)
        if not alive.any(): break
    return total.mean()

def main(seed):
    N, T, E, NB, iters = 64, 128, 10, 32, 40
    oenv = OE.cartpole(); spec = OP.make_spec(4, 2, False)
    params = OP.init_params(spec, seed=seed); st = OU.AdamState.init(params)
    rng = np.random.default_rng(seed + 1)
    y = OE.initial_from_uniform(oenv, rng.random((4, N)).astype(np.float32))
    t = np.zeros(N, np.float32); sc = np.zeros(N, np.int32)
    for it in range(iters):
        na = rng.gumbel(size=(T, N, 2)).astype(np.float32); ru = rng.random((T, N, 4)).astype(np.float32)
        ro = OR.rollout(oenv, spec, params, y, t, sc, T, 0.99, na, ru)
        y, t, sc = ro["y"], ro["t"], ro["step_count"]
        adv, ret = OG.gae(ro["rew"], ro["val"], ro["done"], ro["last_value"], 0.99, 0.95)
        flat = {k: OU.flatten_time_major(v) for k, v in dict(obs=ro["obs"], act=ro["act"], logp=ro["logp"], val=ro["val"], ret=ret, adv=adv).items()}
        idx = np.stack([OU.batch_indices(N * T, N * T // NB, rng) for _ in range(E)])
        params, st, _ = OU.train(spec, params, st, flat, idx, OU.Hyper())
        if it % 10 == 9: print(f"seed {seed} iteration {it + 1}: deterministic return {evaluate(oenv, spec, params):.1f}", flush=True)
    return evaluate(oenv, spec, params)

if __name__ == "__main__":
    r = [main(s) for s in (0, 1, 2)]
    print("final returns:", r, "mean", np.mean(r), "std", np.std(r))
