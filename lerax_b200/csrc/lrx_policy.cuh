// Generic per-lane MLPActorCriticPolicy forward (any widths/depths) and the
// Categorical / Normal arithmetic.  One lane owns one env (or row); activations of a
// warp live in shared memory as [feature][32] columns so reads are conflict-free and the
// weight reads are warp-uniform (broadcast).  Reference:
// policy/actor_critic/mlp.py:133-178, policy/actor.py:40-88,194-261, distreqx 0.0.2.
#pragma once

#include "lrx_common.cuh"

// Layer table in shared/constant form used by the generic kernels.
struct PolicyTable {
  int n_layers_total;
  int first[3];   // index of the first layer of each chain
  int count[3];
  int log_std_off;
  int max_width;  // max over all layer in/out widths
  int feat;       // feature_size
  LrxLayer L[3 * LRX_MAX_LAYERS];
};

inline PolicyTable lrx_make_table(const LrxPolicyDesc& p) {
  PolicyTable t{};
  int np = 0;
  t.n_layers_total = lrx_policy_layers(p, t.L, &t.log_std_off, &np);
  int idx = 0, mw = 1;
  for (int c = 0; c < 3; ++c) {
    t.first[c] = idx;
    t.count[c] = p.n_layers[c];
    idx += p.n_layers[c];
    for (int i = 0; i <= p.n_layers[c]; ++i) mw = p.dims[c][i] > mw ? p.dims[c][i] : mw;
  }
  t.max_width = mw;
  t.feat = p.dims[0][p.n_layers[0]];
  return t;
}

// One Linear (+ReLU) for the lane's column: out[j][lane] = act(b[j] + sum_k W[j][k] in[k][lane]).
// W/b are read through the read-only path with warp-uniform addresses.
__device__ __forceinline__ void lane_linear(const float* __restrict__ params, const LrxLayer& L,
                                            const float* in, float* out, int lane) {
  const float* __restrict__ W = params + L.w_off;
  const float* __restrict__ b = params + L.b_off;
  const int K = L.in, NO = L.out;
  int j0 = 0;
  for (; j0 + 4 <= NO; j0 += 4) {
    float a0 = __ldg(b + j0), a1 = __ldg(b + j0 + 1), a2 = __ldg(b + j0 + 2), a3 = __ldg(b + j0 + 3);
    const float* w0 = W + (size_t)j0 * K;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      float h = in[k * 32 + lane];
      a0 = fmaf(__ldg(w0 + k), h, a0);
      a1 = fmaf(__ldg(w0 + K + k), h, a1);
      a2 = fmaf(__ldg(w0 + 2 * K + k), h, a2);
      a3 = fmaf(__ldg(w0 + 3 * K + k), h, a3);
    }
    if (L.relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f); }
    out[(j0 + 0) * 32 + lane] = a0;
    out[(j0 + 1) * 32 + lane] = a1;
    out[(j0 + 2) * 32 + lane] = a2;
    out[(j0 + 3) * 32 + lane] = a3;
  }
  for (; j0 < NO; ++j0) {
    float a0 = __ldg(b + j0);
    const float* w0 = W + (size_t)j0 * K;
    for (int k = 0; k < K; ++k) a0 = fmaf(__ldg(w0 + k), in[k * 32 + lane], a0);
    if (L.relu) a0 = fmaxf(a0, 0.f);
    out[j0 * 32 + lane] = a0;
  }
}

// Run chain c from `in`; ping-pongs between bufA/bufB (neither may alias `in` unless in == bufA
// or bufB, which is handled) and returns the buffer holding the chain output.
__device__ __forceinline__ float* lane_chain(const float* __restrict__ params, const PolicyTable& T,
                                             int c, float* in, float* bufA, float* bufB, int lane) {
  float* src = in;
  for (int i = 0; i < T.count[c]; ++i) {
    float* dst = (src == bufA) ? bufB : bufA;
    lane_linear(params, T.L[T.first[c] + i], src, dst, lane);
    src = dst;
  }
  return src;
}

// --------------------------------------------------------------------- distributions
// Categorical (distreqx): normalised logits = logits - logsumexp(logits).
template <int MAXN>
__device__ __forceinline__ void log_softmax_n(float (&z)[MAXN], int n) {
  float m = z[0];
#pragma unroll
  for (int i = 1; i < MAXN; ++i) if (i < n) m = fmaxf(m, z[i]);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXN; ++i) if (i < n) s += expf(z[i] - m);
  float lse = logf(s);
#pragma unroll
  for (int i = 0; i < MAXN; ++i) if (i < n) z[i] = (z[i] - m) - lse;
}

#define LRX_HALF_LOG_2PI 0.91893853320467274178f

// Normal (distreqx): log_prob = -0.5*((x-loc)/scale)^2 - (0.5*log(2*pi) + log(scale)).
__device__ __forceinline__ float normal_log_prob(float loc, float scale, float log_scale, float x) {
  float z = (x - loc) / scale;
  return -0.5f * (z * z) - (LRX_HALF_LOG_2PI + log_scale);
}
