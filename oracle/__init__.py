"""CPU oracle for the Lerax on-policy PPO hot path.  TEST INFRASTRUCTURE ONLY.

This package is a NumPy restatement (fp32, with an fp64 twin selected by ``dtype``)
of the reference algorithm for the path named in BASELINE.json.north_star:

    rollout step   -> /root/reference/src/lerax/algorithm/ppo.py:199-316
    envs + Tsit5   -> /root/reference/src/lerax/env/classic_control/*.py, diffrax 0.7.2 (Tsit5)
    policy         -> /root/reference/src/lerax/policy/actor_critic/mlp.py:64-178, policy/actor.py:40-261
    distributions  -> /root/reference/src/lerax/distribution/{categorical,normal}.py -> distreqx 0.0.2
    GAE            -> /root/reference/src/lerax/advantage/gae.py:36-66
    loss / update  -> /root/reference/src/lerax/algorithm/ppo.py:396-561, optax 0.2.6 (chain(clip_by_global_norm, adam))

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it -- as the checker or the timed CPU baseline,
never as a product code path.  ``lerax_b200`` never imports ``oracle``.

PARITY PIN STATUS (see DESIGN.md "Oracle"):
  * GAE and the Categorical/Normal closed forms are PINNED against the reference's
    own known-answer tests (tests/advantage/test_gae.py, tests/distribution/*.py),
    re-expressed in tests/test_oracle_golden.py and tests/golden/*.json.
  * Env dynamics / Tsit5 step, MLP forward, ppo_loss, gradients, clip_by_global_norm,
    Adam and the permutation are "parity unpinned": the reference holds no golden
    vector for them and JAX/diffrax/equinox/optax are not installable in this image
    (no wheels, no network), so the reference cannot be executed here.  They are
    cross-checked instead against independent derivations available in the image:
    Tsit5 order conditions + convergence order vs scipy DOP853, torch autograd (fp64)
    for loss/gradients, torch.optim.Adam + torch clip_grad_norm_ for the optimiser.
"""

from . import philox, tsit5, envs, policy, distributions, gae, rollout, ppo  # noqa: F401
