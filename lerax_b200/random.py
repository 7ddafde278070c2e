"""PRNG keys for the host API (stand-in for `jax.random.key/split`).

The reference threads JAX threefry keys through learn/reset/iteration
(algorithm/base_algorithm.py:231, algorithm/ppo.py:326,353); RNG streams are not part of the
parity contract, so a key here is a 64-bit seed and `split` derives children
with SplitMix64.  The device kernels consume a seed through Philox4x32-10 (include/lerax_b200.h).
"""

from __future__ import annotations

from dataclasses import dataclass

_MASK = (1 << 64) - 1


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _MASK
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
    return z ^ (z >> 31)


@dataclass(frozen=True)
class Key:
    seed: int

    def __int__(self) -> int:
        return self.seed


def key(seed: int) -> Key:
    return Key(_splitmix64(int(seed) & _MASK))


def split(k: Key | int, num: int = 2) -> list[Key]:
    s = int(k) & _MASK
    return [Key(_splitmix64(s ^ _splitmix64(i + 1))) for i in range(num)]
