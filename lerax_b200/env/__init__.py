"""Environment interface of the hot path (reference: env/base_env.py:22-286)."""

from .classic_control import AbstractClassicControlEnv, ClassicControlState

AbstractEnvLike = AbstractClassicControlEnv
AbstractEnv = AbstractClassicControlEnv
AbstractEnvLikeState = ClassicControlState
AbstractEnvState = ClassicControlState

__all__ = ["AbstractEnvLike", "AbstractEnv", "AbstractEnvLikeState", "AbstractEnvState"]
