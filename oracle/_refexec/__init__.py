"""oracle/_refexec -- runs the UNMODIFIED reference source of the on-policy hot path without JAX.

TEST INFRASTRUCTURE.  Nothing under lerax_b200/ imports this; only tests/golden/make_refexec.py (the fixture
generator, run in the authoring container where /root/reference exists) does.

The reference (RunnersNum40/lerax 0.0.6) is pure Python on jax / equinox / diffrax / distreqx / optax, none of which
is installable here.  `shim/` holds small stand-ins for exactly the third-party calls the hot path makes, on torch CPU
fp32 tensors (so `eqx.filter_value_and_grad` can use torch autograd); `loader.load()` puts them on sys.path, registers
`lerax` as a namespace over /root/reference/src/lerax (package __init__ files that drag in rendering / MuJoCo / logging
back ends are skipped) and imports the reference's own modules.  What then executes is the reference's code, line for
line: env `initial / dynamics / transition / clip / reward / terminal / observation`, `TimeLimit`, `GAE.__call__`,
`MLPActorCriticPolicy.action_and_value / value / evaluate_action`, `PPO.step`, `collect_rollout`, `ppo_loss`,
`train_batch`, `train_epoch`, `train`, `RolloutBuffer.flatten_axes / batch_indices / gather`.
Third-party arithmetic restated in the stand-ins (each cites its pinned version): one explicit Tsit5 step (diffrax
0.7.2), Linear / MLP (equinox 0.13.6), Categorical / Normal (distreqx 0.0.2), clip_by_global_norm + adam (optax 0.2.6),
jnp / lax semantics (jax 0.9.0.1).  Random draws are injected by key path (RNG streams are outside the parity contract).
"""
