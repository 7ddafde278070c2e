"""lerax_b200 -- B200-native (sm_100a CUDA) backend for Lerax's on-policy PPO hot path.

Drop-in for `PPO().learn(env, policy, total_timesteps, key=...)` and the env / policy /
buffer / advantage / callback interfaces on that path (reference: RunnersNum40/lerax 0.0.6).
Host code is Python; all arithmetic runs in hand-written CUDA kernels behind the C ABI in
include/lerax_b200.h (lerax_b200/_C/liblerax_b200.so).  No CPU fallback.
"""

__version__ = "0.1.0"

from . import random  # noqa: F401
