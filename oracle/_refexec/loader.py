"""Import the reference's hot-path modules over the stand-ins in shim/ (see package docstring)."""

from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_SRC = os.environ.get("LERAX_REFERENCE_SRC", "/root/reference/src")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")

# packages whose __init__.py would import rendering / physics / logging back ends: registered as bare namespaces,
# their public names resolved lazily from the listed submodules
_NAMESPACES = {
    "lerax": [],
    "lerax.env": ["base_env"],
    "lerax.env.classic_control": ["acrobot", "cartpole", "continuous_mountain_car", "mountain_car", "pendulum"],
    "lerax.callback": ["base_callback", "empty", "list"],
    "lerax.policy": ["base_policy", "actor_critic"],
    "lerax.wrapper": ["base_wrapper", "misc"],
    "lerax.algorithm": ["base_algorithm", "ppo", "a2c", "reinforce"],
    "lerax.buffer": ["base_buffer", "rollout"],
}


class _Lenient(types.ModuleType):
    """A module whose every attribute is a dummy class (lerax.render: renderers are never constructed here)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (), {})
        setattr(self, name, cls)
        return cls


class _Namespace(types.ModuleType):
    def __init__(self, name, path, submodules):
        super().__init__(name)
        self.__path__ = [path]
        self.__package__ = name
        self._submodules = submodules

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        for sub in self._submodules:
            mod = importlib.import_module(f"{self.__name__}.{sub}")
            if hasattr(mod, name):
                return getattr(mod, name)
        try:
            return importlib.import_module(f"{self.__name__}.{name}")
        except ImportError as e:
            raise AttributeError(f"{self.__name__}.{name}: {e}") from e


_loaded = {}


def load():
    """Returns a namespace object with the reference classes used by the fixture generator."""
    if _loaded:
        return _loaded["ns"]
    if not os.path.isdir(os.path.join(REFERENCE_SRC, "lerax")):
        raise FileNotFoundError(f"reference source not found under {REFERENCE_SRC}")
    for m in ("jax", "equinox", "diffrax", "distreqx", "optax", "lerax"):
        if m in sys.modules and not getattr(sys.modules[m], "__file__", "" or "").startswith(_SHIM):
            raise RuntimeError(f"{m} is already imported from elsewhere")
    sys.path.insert(0, _SHIM)
    sys.modules["lerax.render"] = _Lenient("lerax.render")
    for name, subs in _NAMESPACES.items():
        rel = name.split(".")[1:]
        sys.modules[name] = _Namespace(name, os.path.join(REFERENCE_SRC, "lerax", *rel), subs)
    for name in _NAMESPACES:  # parent.child attribute links
        if "." in name:
            parent, child = name.rsplit(".", 1)
            object.__setattr__(sys.modules[parent], child, sys.modules[name])
    sys.modules["lerax"].render = sys.modules["lerax.render"]

    ns = types.SimpleNamespace()
    import jax
    import jax.numpy as jnp
    import jax.random as jr

    ns.jax, ns.jnp, ns.jr = jax, jnp, jr
    ns.eqx = importlib.import_module("equinox")
    ns.optax = importlib.import_module("optax")
    cc = "lerax.env.classic_control."
    ns.CartPole = importlib.import_module(cc + "cartpole").CartPole
    ns.Pendulum = importlib.import_module(cc + "pendulum").Pendulum
    ns.Acrobot = importlib.import_module(cc + "acrobot").Acrobot
    ns.MountainCar = importlib.import_module(cc + "mountain_car").MountainCar
    ns.ContinuousMountainCar = importlib.import_module(cc + "continuous_mountain_car").ContinuousMountainCar
    ns.TimeLimit = importlib.import_module("lerax.wrapper.misc").TimeLimit
    ns.GAE = importlib.import_module("lerax.advantage.gae").GAE
    ns.NStepReturn = importlib.import_module("lerax.advantage.n_step").NStepReturn
    ns.BootstrappedReturn = importlib.import_module("lerax.advantage.bootstrapped").BootstrappedReturn
    ns.RolloutBuffer = importlib.import_module("lerax.buffer.rollout").RolloutBuffer
    ns.MLPActorCriticPolicy = importlib.import_module("lerax.policy.actor_critic.mlp").MLPActorCriticPolicy
    ppo = importlib.import_module("lerax.algorithm.ppo")
    ns.PPO, ns.PPOStepState = ppo.PPO, ppo.PPOStepState
    ns.A2C = importlib.import_module("lerax.algorithm.a2c").A2C
    ns.REINFORCE = importlib.import_module("lerax.algorithm.reinforce").REINFORCE
    ns.CallbackList = importlib.import_module("lerax.callback.list").CallbackList
    ns.Box = importlib.import_module("lerax.space.box").Box
    _loaded["ns"] = ns
    return ns
