"""Philox4x32-10 counter-based RNG + the noise transforms the CUDA rollout uses.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference draws its noise from
JAX threefry keys (/root/reference/src/lerax/algorithm/ppo.py:209-219); RNG streams
are not part of the parity contract -- only the sampling
distributions are.  The CUDA path uses Philox4x32-10 (Salmon et al., SC'11, Random123)
keyed as documented below; this file restates it so free-running rollouts are checkable.

Keying:  key = (seed & 0xffffffff, seed >> 32)
         counter = (global_env_index, global_step_index, stream, block)
         stream 0 = action noise, stream 1 = reset noise.
"""

from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)

STREAM_ACTION = 0
STREAM_RESET = 1


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint32) for x in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for r in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & mask).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & mask).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            if r < 9:
                k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
                k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def seed_to_key(seed: int):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, seed >> 32


def u32_to_unit(bits, dtype=np.float32):
    """(bits >> 9 + 0.5) * 2^-23 : a float in the OPEN interval (0, 1), exact in fp32."""
    x = (np.asarray(bits, dtype=np.uint32) >> np.uint32(9)).astype(dtype)
    return x * dtype(2.0**-23) + dtype(2.0**-24)


def raw_words(seed, env_idx, step_idx, stream, block=0):
    k0, k1 = seed_to_key(seed)
    return philox4x32_10(env_idx, step_idx, np.uint32(stream), np.uint32(block), k0, k1)


def uniforms(seed, env_idx, step_idx, stream, n, dtype=np.float32):
    """n (<=4) uniforms in (0,1) per (env, step); returns array [..., n]."""
    assert 1 <= n <= 4
    w = raw_words(seed, env_idx, step_idx, stream)
    return np.stack([u32_to_unit(w[i], dtype) for i in range(n)], axis=-1)


def gumbels(seed, env_idx, step_idx, n, dtype=np.float32):
    """Gumbel(0,1) noise for Categorical sampling: g = -log(-log(u))."""
    u = uniforms(seed, env_idx, step_idx, STREAM_ACTION, n, dtype)
    return -np.log(-np.log(u))


def normals(seed, env_idx, step_idx, dtype=np.float32):
    """One N(0,1) draw per (env, step) by Box-Muller on words 0,1."""
    u = uniforms(seed, env_idx, step_idx, STREAM_ACTION, 2, dtype)
    r = np.sqrt(dtype(-2.0) * np.log(u[..., 0]))
    return r * np.cos(dtype(2.0 * np.pi) * u[..., 1])


def reset_uniforms(seed, env_idx, step_idx, n, dtype=np.float32):
    return uniforms(seed, env_idx, step_idx, STREAM_RESET, n, dtype)
