// Shared host/device helpers for the lerax_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/lerax_b200.h"

// ----------------------------------------------------------------------------- errors
void lrx_set_error(const char* fmt, ...);
void lrx_count_launch(int n = 1);

#define LRX_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      lrx_set_error(__VA_ARGS__);     \
      return (code);                  \
    }                                 \
  } while (0)

#define LRX_CUDA_CHECK(expr)                                                        \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      lrx_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                    __LINE__);                                                      \
      return LRX_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define LRX_LAUNCH_CHECK()                 \
  do {                                     \
    lrx_count_launch();                    \
    LRX_CUDA_CHECK(cudaGetLastError());    \
  } while (0)

constexpr int LRX_MAX_SMS = 256;  // upper bound used to SIZE per-SM workspaces (B200 has 148)
int lrx_sm_count();               // multiprocessors of the CURRENT device (cudaDeviceGetAttribute)
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per device and kernel: `done` is a per-kernel table indexed by
// the device ordinal (an idempotent cache, not state the results depend on)
template <typename K>
inline cudaError_t lrx_set_max_smem(K kernel, int bytes, bool (&done)[64]) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
  return e;
}

// ------------------------------------------------------------------- policy descriptor
struct LrxLayer {
  int chain, layer, w_off, b_off, in, out, relu;
};

// Flat (Equinox leaf order) offsets; returns number of Linear layers written to `out`
// (at most 3*LRX_MAX_LAYERS) and the log_std offset (-1 for Discrete) / total params.
__host__ __device__ inline int lrx_policy_layers(const LrxPolicyDesc& p, LrxLayer* out,
                                                 int* log_std_off, int* n_params) {
  int off = 0, n = 0;
  for (int c = 0; c < 3; ++c) {
    for (int i = 0; i < p.n_layers[c]; ++i) {
      int k = p.dims[c][i], o = p.dims[c][i + 1];
      if (out) out[n] = LrxLayer{c, i, off, off + k * o, k, o, p.relu[c][i]};
      ++n;
      off += k * o + o;
    }
  }
  if (log_std_off) *log_std_off = p.is_box ? off : -1;
  if (p.is_box) off += 1;
  if (n_params) *n_params = off;
  return n;
}

int lrx_fused_enabled();
int lrx_pdl_enabled();
int lrx_validate_policy(const LrxPolicyDesc* pol);
int lrx_validate_env(const LrxEnvDesc* env);
int lrx_env_state_dim(int kind);
int lrx_env_obs_dim(int kind);
bool lrx_env_is_box(int kind);
int lrx_env_num_outputs(int kind);

// --------------------------------------------------------------------------- Philox
// Philox4x32-10 (Salmon et al., SC'11).  Restated in oracle/philox.py; KAT in tests/golden.
struct Philox4 {
  uint32_t x, y, z, w;
};

__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                 uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0;
    uint64_t p1 = (uint64_t)M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

constexpr uint32_t LRX_STREAM_ACTION = 0;
constexpr uint32_t LRX_STREAM_RESET = 1;

// (bits >> 9 + 0.5) * 2^-23: open interval (0,1), exact in fp32.
__host__ __device__ inline float lrx_u32_to_unit(uint32_t bits) {
  return (float)(bits >> 9) * 1.1920928955078125e-07f + 5.9604644775390625e-08f;
}

__device__ inline Philox4 lrx_noise_words(const LrxRng& rng, uint32_t env, uint32_t step,
                                          uint32_t stream) {
  return philox4x32_10(rng.env_offset + env, rng.step_offset + step, stream, 0u,
                       (uint32_t)rng.seed, (uint32_t)(rng.seed >> 32));
}

__device__ __forceinline__ float lrx_clipf(float x, float lo, float hi) {
  return fminf(fmaxf(x, lo), hi);  // jnp.clip = minimum(maximum(x, lo), hi)
}

// --------------------------------------------------------------------- warp helpers
__device__ inline float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ inline double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
