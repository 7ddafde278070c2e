// PPO update: minibatch gather, MLP forward/backward, clipped-surrogate / value / entropy
// loss, gradient reduction, clip_by_global_norm and Adam.
//
// Reference: algorithm/ppo.py:396-467 (ppo_loss), :471-492 (train_batch), :494-521
// (train_epoch); buffer/base_buffer.py:38-103 (flatten i = n*T + t, gather); optax 0.2.6
// chain(clip_by_global_norm, adam) (ppo.py:192-194).
//
// This file holds the GENERIC path (any layer widths/depths): activations of a row chunk
// are kept feature-major [width][rows] in the caller's workspace and every Linear is one
// tiled fp32 SGEMM launch (forward, dX, and split-K dW with the bias gradient folded in as a
// ones column).  The fused per-warp path for the default widths lives in
// lrx_update_fused.cu and is tried first.
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#include "lrx_common.cuh"
#include "lrx_gemm.cuh"
#include "lrx_p2p.cuh"
#include "lrx_policy.cuh"

constexpr int64_t LRX_CHUNK_ROWS = 1 << 20;  // rows of a minibatch processed per pass
constexpr int LRX_LOSS_THREADS = 256;
constexpr int LRX_MAX_SPLIT = 128;

// ============================================================================== SGEMM
// C[M x N] (+)= op(A)[M x K] * op(B)[K x N], fp32, 64x64x16 tiles, 4x4 micro-tiles.
template <bool AT, bool BT>
__global__ void __launch_bounds__(256) sgemm_kernel(const GemmArgs g) {
  constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;
  __shared__ float As[BK][BM + PAD];
  __shared__ float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * g.k_chunk;
  const int kend = min(g.K, kbeg + g.k_chunk);
  float acc[4][4] = {};

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- A tile -> As[k][i]
    if (AT) {
      const int k = k0 + (tid >> 4), ib = i0 + (tid & 15) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = ib + q;
        As[tid >> 4][(tid & 15) * 4 + q] = (k < kend && i < g.M) ? g.A[(size_t)k * g.lda + i] : 0.f;
      }
    } else {
      const int i = i0 + (tid >> 2), kb = k0 + (tid & 3) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = kb + q;
        As[(tid & 3) * 4 + q][tid >> 2] = (k < kend && i < g.M) ? g.A[(size_t)i * g.lda + k] : 0.f;
      }
    }
    // ---- B tile -> Bs[k][j]
    if (BT) {
      const int j = j0 + (tid >> 2), kb = k0 + (tid & 3) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = kb + q;
        Bs[(tid & 3) * 4 + q][tid >> 2] = (k < kend && j < g.N) ? g.B[(size_t)j * g.ldb + k] : 0.f;
      }
    } else {
      const int k = k0 + (tid >> 4), jb = j0 + (tid & 15) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = jb + q;
        Bs[tid >> 4][(tid & 15) * 4 + q] = (k < kend && j < g.N) ? g.B[(size_t)k * g.ldb + j] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(av[p], bv[q], acc[p][q]);
    }
    __syncthreads();
  }

  float* C = g.C + (size_t)blockIdx.z * g.c_split_stride;
  float* C2 = g.C2 ? g.C2 + (size_t)blockIdx.z * g.c_split_stride : nullptr;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int i = i0 + ty * 4 + p;
    if (i >= g.M) continue;
    const float bi = g.bias ? g.bias[i] : 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = j0 + tx * 4 + q;
      if (j >= g.N) continue;
      float v = acc[p][q] + bi;
      if (g.relu) v = fmaxf(v, 0.f);
      if (g.mask) v = (g.mask[(size_t)i * g.ldm + j] > 0.f) ? v : 0.f;
      float* dst = (C2 && j == g.n_main) ? (C2 + i) : (C + (size_t)i * g.ldc + j);
      if (g.accumulate) v += *dst;
      *dst = v;
    }
  }
  if (g.ones_row && blockIdx.y == 0 && blockIdx.z == 0) {
    for (int j = j0 + tid; j < min(g.N, j0 + BN); j += 256) g.C[(size_t)g.M * g.ldc + j] = 1.0f;
  }
}

static int launch_gemm(const GemmArgs& g, bool at, bool bt, int splits, cudaStream_t s) {
  bool handled = false;  // layer widths and contractions of at least 16: tensor cores (lrx_gemm_tc.cu)
  const int rc = lrx_launch_gemm_tc(g, at, bt, splits, s, &handled);
  if (rc || handled) return rc;
  dim3 grid((g.N + 63) / 64, (g.M + 63) / 64, splits);
  if (!at && !bt) sgemm_kernel<false, false><<<grid, 256, 0, s>>>(g);
  else if (at && !bt) sgemm_kernel<true, false><<<grid, 256, 0, s>>>(g);
  else if (!at && bt) sgemm_kernel<false, true><<<grid, 256, 0, s>>>(g);
  else sgemm_kernel<true, true><<<grid, 256, 0, s>>>(g);
  LRX_LAUNCH_CHECK();
  return LRX_OK;
}

// exact-fp32 GEMM for callers outside this file (the step-by-step rollout of wide policies, lrx_rollout_steps.cu)
int lrx_launch_gemm_fma(const GemmArgs& g, bool at, bool bt, int splits, cudaStream_t s) {
  dim3 grid((g.N + 63) / 64, (g.M + 63) / 64, splits);
  if (!at && !bt) sgemm_kernel<false, false><<<grid, 256, 0, s>>>(g);
  else if (at && !bt) sgemm_kernel<true, false><<<grid, 256, 0, s>>>(g);
  else if (!at && bt) sgemm_kernel<false, true><<<grid, 256, 0, s>>>(g);
  else sgemm_kernel<true, true><<<grid, 256, 0, s>>>(g);
  LRX_LAUNCH_CHECK();
  return LRX_OK;
}

// The layer-by-layer update's GEMM by itself (tools/gemm_time.py, tests): C[M x N] (+)= op(A) op(B) with the epilogue
// options of GemmArgs; use_tc = 0 forces the FMA kernel.
extern "C" int lrx_debug_gemm(const float* A, int32_t lda, const float* B, int32_t ldb, float* C, int32_t ldc, int32_t M,
                              int32_t N, int32_t K, int32_t at, int32_t bt, const float* bias, int32_t relu, const float* mask,
                              int32_t ldm, int32_t accumulate, int32_t splits, int32_t k_chunk, int64_t c_split_stride,
                              int32_t use_tc, lrx_stream_t stream) {
  LRX_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && splits >= 1, LRX_ERR_INVALID_ARG, "lrx_debug_gemm: bad argument");
  GemmArgs g{};
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K; g.bias = bias; g.relu = relu; g.mask = mask; g.ldm = ldm; g.accumulate = accumulate;
  g.k_chunk = k_chunk > 0 ? k_chunk : K; g.c_split_stride = c_split_stride; g.n_main = -1;
  cudaStream_t s = (cudaStream_t)stream;
  return use_tc ? launch_gemm(g, at != 0, bt != 0, splits, s) : lrx_launch_gemm_fma(g, at != 0, bt != 0, splits, s);
}

// ============================================================================= gather
// Rows idx[m] (reference order i = n*T + t) -> feature-major obsT [(d+1)][rows] with a ones
// row, plus the per-row scalars.  buffer/base_buffer.py:90-103 (gather).
__global__ void gather_kernel(LrxBatchView buf, const int32_t* __restrict__ idx, int rows, int ld,
                              int d, int is_box, float* obsT, float* act_f, int32_t* act_i,
                              float* logp, float* adv, float* ret, float* val) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= rows) return;
  const int i = idx[m];
  const int n = i / buf.T, t = i - n * buf.T;
  const size_t src = (size_t)t * buf.N + n;
  for (int k = 0; k < d; ++k) obsT[(size_t)k * ld + m] = buf.obs[src * d + k];
  obsT[(size_t)d * ld + m] = 1.0f;
  if (is_box) act_f[m] = ((const float*)buf.act)[src];
  else act_i[m] = ((const int32_t*)buf.act)[src];
  logp[m] = buf.logp[src];
  adv[m] = buf.adv[src];
  ret[m] = buf.ret[src];
  val[m] = buf.val[src];
}

// ========================================================================== adv sums
// One thread-block CLUSTER per minibatch (up to 8 CTAs, so 32 minibatches fill the GPU instead of 32 SMs): every thread a
// fixed strided subset of the rows, fixed-order tree inside the CTA, then CTA 0 of the cluster adds the CTA totals in rank
// order through distributed shared memory.  The sums -- and with them the normalised advantages and every gradient -- are
// bitwise reproducible (no atomics, no memset, no scratch).
__global__ void __launch_bounds__(1024)
adv_sums_kernel(const float* __restrict__ adv, int N, int T, const int32_t* __restrict__ idx, long long B, double* sums) {
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned S = cluster.num_blocks(), rank = cluster.block_rank();
  const int b = blockIdx.x / S;
  const int32_t* row = idx + (size_t)b * B;
  double s = 0, s2 = 0;
  constexpr int U = 8;  // gathered loads in flight per thread
  for (long long m0 = (long long)rank * U * blockDim.x + threadIdx.x; m0 < B; m0 += (long long)U * blockDim.x * S) {
    float a[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long m = m0 + (long long)u * blockDim.x;
      a[u] = 0.f;
      if (m < B) {
        const int i = __ldg(row + m);
        const int n = i / T, t = i - n * T;
        a[u] = __ldg(adv + (size_t)t * N + n);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      s += a[u];
      s2 += (double)a[u] * a[u];
    }
  }
  s = warp_sum(s);
  s2 = warp_sum(s2);
  __shared__ double sh[2][32];
  __shared__ double tot[2];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sh[0][w] = s; sh[1][w] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double x = 0, c = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { x += sh[0][i]; c += sh[1][i]; }
    tot[0] = x;
    tot[1] = c;
  }
  cluster.sync();
  if (rank == 0 && threadIdx.x == 0) {
    double x = 0, c = 0;
    for (unsigned k = 0; k < S; ++k) {
      const double* peer = cluster.map_shared_rank(tot, k);
      x += peer[0];
      c += peer[1];
    }
    sums[2 * b] = x;
    sums[2 * b + 1] = c;
  }
  cluster.sync();  // the peers' shared memory must outlive CTA 0's reads
}

// =============================================================================== loss
// Per row: log-prob / entropy of the stored action under the current policy, ratio,
// normalised advantage, clipped surrogate, value loss; emits dL/dvalue and dL/dhead (already
// scaled by 1/B_global) and block partial sums of the statistics.   ppo.py:396-467.
struct LossArgs {
  const float* value;  // [rows]
  const float* head; int ld;  // [n_out][ld]
  const float* act_f; const int32_t* act_i;
  const float* logp_old; const float* adv; const float* ret; const float* val_old;
  float* dvalue; float* dhead;  // [rows], [n_out][ld]
  int rows; int n_out; int is_box;
  const float* log_std_ptr;  // device scalar (Box only)
  const double* adv_sums; double B_global;
  LrxPpoHyper h;
  double* partial;  // [gridDim.x][8]
};

__device__ __forceinline__ void tie_weights(float a, float b, float& wa, float& wb) {
  wa = a < b ? 1.f : (a == b ? 0.5f : 0.f);  // JAX: d min(a,b) splits ties 1/2-1/2
  wb = 1.f - wa;
}
__device__ __forceinline__ float clip_grad(float x, float lo, float hi) {
  return (x > lo && x < hi) ? 1.f : ((x == lo || x == hi) ? 0.5f : 0.f);
}

__global__ void __launch_bounds__(LRX_LOSS_THREADS) loss_kernel(const LossArgs a) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const float invB = (float)(1.0 / a.B_global);
  double s_kl = 0, s_pol = 0, s_val = 0, s_ent = 0, s_dls = 0, s_bad = 0;
  if (m < a.rows) {
    const float value = a.value[m];
    float logp, entropy, z_n = 0.f, scale = 1.f;
    float lp[LRX_MAX_NOUT], pr[LRX_MAX_NOUT];
    int ai = 0;
    if (a.is_box) {
      scale = expf(__ldg(a.log_std_ptr));
      const float loc = a.head[m];
      z_n = (a.act_f[m] - loc) / scale;
      const float log_scale = logf(scale);
      logp = -0.5f * (z_n * z_n) - (LRX_HALF_LOG_2PI + log_scale);
      entropy = 0.5f + LRX_HALF_LOG_2PI + log_scale;
    } else {
#pragma unroll
      for (int j = 0; j < LRX_MAX_NOUT; ++j) lp[j] = j < a.n_out ? a.head[(size_t)j * a.ld + m] : 0.f;
      log_softmax_n<LRX_MAX_NOUT>(lp, a.n_out);
      ai = a.act_i[m];
      logp = 0.f;
      entropy = 0.f;
#pragma unroll
      for (int j = 0; j < LRX_MAX_NOUT; ++j) {
        if (j < a.n_out) {
          pr[j] = expf(lp[j]);
          entropy -= pr[j] * lp[j];
          if (j == ai) logp = lp[j];
        }
      }
    }
    if (!(isfinite(value) && isfinite(logp) && isfinite(entropy))) s_bad = 1.0;  // ppo.py:413-417

    const float log_ratio = logp - a.logp_old[m];
    const float ratio = expf(log_ratio);
    s_kl = (double)(ratio - log_ratio);

    float A = a.adv[m];
    if (a.h.normalize_advantages) {
      const double mean = a.adv_sums[0] / a.B_global;
      double var = a.adv_sums[1] / a.B_global - mean * mean;
      var = var > 0 ? var : 0;
      const float std = (float)sqrt(var);
      A = (A - (float)mean) / (std + 1.1920928955078125e-07f);  // + finfo(float32).eps
    }
    const float c = a.h.clip_coefficient, lo = 1.f - c, hi = 1.f + c;
    float dlogp;
    if (a.h.surrogate == 1) {  // A2C / REINFORCE: -mean(logp * A)   (a2c.py:401, reinforce.py:391)
      s_pol = (double)(logp * A);
      dlogp = -invB * A;
    } else {
      const float s1 = A * ratio, s2 = A * lrx_clipf(ratio, lo, hi);
      s_pol = (double)fminf(s1, s2);
      float w1, w2;
      tie_weights(s1, s2, w1, w2);
      dlogp = -invB * (w1 * A * ratio + w2 * A * clip_grad(ratio, lo, hi) * ratio);
    }

    float dvl;
    const float e1 = value - a.ret[m];
    if (a.h.clip_value_loss) {
      const float dv = value - a.val_old[m];
      const float clipped = a.val_old[m] + lrx_clipf(dv, -c, c);
      const float e2 = clipped - a.ret[m];
      const float q1 = e1 * e1, q2 = e2 * e2;
      s_val = (double)fminf(q1, q2);
      float u1, u2;
      tie_weights(q1, q2, u1, u2);
      dvl = invB * (u1 * e1 + u2 * e2 * clip_grad(dv, -c, c));
    } else {
      s_val = (double)(e1 * e1);
      dvl = invB * e1;
    }
    s_ent = (double)entropy;
    a.dvalue[m] = a.h.value_loss_coefficient * dvl;
    const float ch = a.h.entropy_loss_coefficient;
    if (a.is_box) {
      a.dhead[m] = dlogp * z_n / scale;
      s_dls = (double)(dlogp * (z_n * z_n - 1.0f));
    } else {
#pragma unroll
      for (int j = 0; j < LRX_MAX_NOUT; ++j) {
        if (j < a.n_out) {
          const float dH = -pr[j] * (lp[j] + entropy);
          a.dhead[(size_t)j * a.ld + m] = dlogp * ((j == ai ? 1.f : 0.f) - pr[j]) + (-ch * invB) * dH;
        }
      }
    }
  }
  // block partial sums (fixed order -> deterministic)
  __shared__ double sh[6][LRX_LOSS_THREADS / 32];
  double v[6] = {s_kl, s_pol, s_val, s_ent, s_dls, s_bad};
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    v[q] = warp_sum(v[q]);
    if (lane == 0) sh[q][w] = v[q];
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double s = 0;
    for (int i = 0; i < LRX_LOSS_THREADS / 32; ++i) s += sh[threadIdx.x][i];
    a.partial[(size_t)blockIdx.x * 8 + threadIdx.x] = s;
  }
}

// =========================================================== partial-gradient reduction
// gradbuf[p] = sum_z partial[z][p]; the last block also folds the loss partial sums into
// the log_std gradient and the stat slots.  Fixed summation order -> deterministic.
// Block = 32 parameters x 8 row groups: each thread sums every 8th partial row of one parameter
// (independent loads in flight), the 8 sub-sums are combined in a fixed order.  With `ap.params`
// set (single-GPU update) the kernel also applies the update: every block publishes the sum of
// squares of its 32 gradient entries; after a grid barrier every block adds them in block order ->
// global norm -> clip + Adam for its own 32 parameters.  Deterministic.
struct LrxApplyArgs {
  float* params; float* m; float* v; int32_t* count; LrxPpoHyper h; float* stats_out; int32_t* nonfinite_flag;
  double* blk_ss;        // [gridDim.x]
  // multi-GPU (lrx_ppo_grad_apply_p2p): `gradbuf` is this rank's exchange slot; after the local reduction every block
  // publishes ITS 32 entries (per-block flag), waits for the same block of every peer and sums the G slots in rank order
  char* const* peers;    // device array of the G exchange buffers; nullptr = single GPU
  int rank, world;
  size_t slot_b;
  uint32_t seq;
  int* error_flag;
  float* gsum;           // [P + LRX_PPO_NSTATS]: the summed statistics and log_std entry (block 0 -> everyone, after the grid barrier)
  // multi-GPU, three-call path (lrx_ppo_grad_p2p; params == nullptr): the LAST block to finish the reduction publishes the
  // slot (own global flag = publish_seq), so the flag travels to the peers while lrx_ppo_apply_p2p is still being launched
  uint32_t* publish_flag;
  uint32_t publish_seq;
  unsigned* ticket;      // zero between launches (the last block resets it)
  // ... or PUSHES it (push != 0; peers / rank / world / slot_b above): every entry is also written into each rank's receive
  // area of this rank, and the last block sets this rank's arrival flag in every buffer
  int push;
};

__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ partial, int splits, long long stride, int P,
                       const double* __restrict__ loss_partial, int loss_blocks, int log_std_off,
                       float ent_coef, double B_global, float* gradbuf, const LrxApplyArgs ap) {
  __shared__ float sub[8][33];
  __shared__ double red[8];
  // programmatic dependent launch (see ppo_grad_tc128c_kernel): the partials are the previous kernel's output
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int pl = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int p = blockIdx.x * 32 + pl;
  float s = 0.f;
  if (p < P) {
#pragma unroll 4
    for (int z = rg; z < splits; z += 8) s += partial[(size_t)z * stride + p];
  }
  sub[rg][pl] = s;
  __syncthreads();
  float t = 0.f;
  if (rg == 0 && p < P && p != log_std_off) {
    t = sub[0][pl];
#pragma unroll
    for (int k = 1; k < 8; ++k) t += sub[k][pl];
    gradbuf[p] = t;
  }
  // loss statistics: warp q (< 6) of block 0 reduces quantity q; lane-strided sums then a
  // fixed-order shuffle tree
  double lsg = 0;  // log_std gradient (block 0, warp 4, lane 0)
  if (blockIdx.x == 0 && rg < 6) {
    double d = 0;
    for (int b = pl; b < loss_blocks; b += 32) d += loss_partial[(size_t)b * 8 + rg];
    d = warp_sum(d);
    if (pl == 0) {
      if (rg < 4) gradbuf[P + rg] = (float)(d / B_global);       // kl, pol, val, ent means' share
      else if (rg == 4) { if (log_std_off >= 0) { gradbuf[log_std_off] = (float)d; lsg = (double)(float)d; } }
      else gradbuf[P + 6] = (float)d;                              // non-finite row count
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 255) { gradbuf[P + 4] = 0.f; gradbuf[P + 5] = 0.f; gradbuf[P + 7] = 0.f; }
  if (ap.publish_flag) {
    if (ap.push) {
      __syncthreads();  // block 0: the statistics and the log_std entry written above
      // this block's entries of the own slot -> receive area [slot][this rank] of EVERY rank (its own included)
      const int n_mine = blockIdx.x == 0 ? 32 + LRX_PPO_NSTATS + 1 : 32;
      if ((int)threadIdx.x < n_mine) {
        int e = -1;
        if (threadIdx.x < 32) e = (p < P) ? p : -1;
        else if (threadIdx.x < 32 + LRX_PPO_NSTATS) e = P + ((int)threadIdx.x - 32);
        else e = log_std_off;  // -1 for Discrete policies
        if (e >= 0) {
          const float val = (threadIdx.x < 32 && p != log_std_off) ? t : gradbuf[e];
          if (threadIdx.x >= 32 || p != log_std_off)
            for (int r = 0; r < ap.world; ++r) p2p::push_recv(ap.peers[r], ap.slot_b, ap.publish_seq, ap.rank)[e] = val;
        }
      }
      __threadfence_system();
    } else {
      __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(ap.ticket, 1u) == gridDim.x - 1) {
      *ap.ticket = 0u;
      __threadfence_system();  // every block's stores (observed through the ticket) before the flag(s)
      if (ap.push) {
        for (int r = 0; r < ap.world; ++r) p2p::st_release_sys(p2p::push_flag(ap.peers[r], ap.rank), ap.publish_seq);
      } else {
        p2p::st_release_sys(ap.publish_flag, ap.publish_seq);
      }
    }
  }
  if (!ap.params) return;

  // ---- multi-GPU: the exchange, block by block.  Block b's 32 entries of the own slot are complete -> release its flag;
  // wait for block b of every peer (they run the same grid); sum the G slots in rank order (bit-identical on every rank).
  // No grid-wide barrier and no second launch between the local reduction and the exchange; two slots suffice because a
  // rank reaches seq + 2 only after its block b saw every peer's block b at seq + 1, i.e. after they finished kernel seq.
  const bool p2p_on = ap.peers != nullptr;
  bool peer_lost = false;
  double lsg_ss = 0;  // p2p: the summed log_std entry's contribution to the norm (block 0, warp 1, lane 8)
  if (p2p_on) {
    // LRX_P2P_FUSED=2 (ap.push == 2 here): ONE flag per rank, published by the last block to finish the reduction (ticket),
    // instead of one per block: a single system-scope fence and flag store per rank and minibatch
    // LRX_P2P_FUSED=3 (ap.push == 3): as 2, but PUSH -- the entries also go into every rank's receive area of this rank and
    // the last block sets this rank's arrival flag in every buffer; receivers poll and sum local memory only
    const bool fpush = ap.push == 3;
    const bool one_flag = ap.push == 2 || fpush;
    if (fpush) {
      __syncthreads();  // block 0: the statistics and the log_std entry written above
      const int n_mine = blockIdx.x == 0 ? 32 + LRX_PPO_NSTATS + 1 : 32;
      if ((int)threadIdx.x < n_mine) {
        int e = -1;
        if (threadIdx.x < 32) e = (p < P && p != log_std_off) ? p : -1;
        else if (threadIdx.x < 32 + LRX_PPO_NSTATS) e = P + ((int)threadIdx.x - 32);
        else e = log_std_off;  // -1 for Discrete policies
        if (e >= 0) {
          const float val = threadIdx.x < 32 ? t : gradbuf[e];
          for (int r = 0; r < ap.world; ++r) p2p::push_recv(ap.peers[r], ap.slot_b, ap.seq, ap.rank)[e] = val;
        }
      }
      __threadfence_system();
    } else if (one_flag) {
      __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (one_flag) {
        if (atomicAdd(ap.ticket, 1u) == gridDim.x - 1) {
          *ap.ticket = 0u;
          __threadfence_system();
          if (fpush) {
            for (int r = 0; r < ap.world; ++r) p2p::st_release_sys(p2p::push_flag(ap.peers[r], ap.rank), ap.seq);
          } else {
            p2p::st_release_sys(reinterpret_cast<uint32_t*>(ap.peers[ap.rank]), ap.seq);
          }
        }
      } else {
        __threadfence_system();
        p2p::st_release_sys(p2p::block_flag(ap.peers[ap.rank], ap.seq, blockIdx.x), ap.seq);
      }
    }
    int timed_out = 0;
    // one_flag: only block 0 polls the peers' flags over NVLink (G threads per rank; every block polling them was 218 us
    // per minibatch at 8 GPUs against 201 us for the three-call path) and relays "everyone has arrived" through a flag in
    // LOCAL memory (own buffer, offset 72) that the other blocks poll
    uint32_t* relay = reinterpret_cast<uint32_t*>(ap.peers[ap.rank] + 72);
    const bool remote_poller = !one_flag || blockIdx.x == 0;
    if (remote_poller && (int)threadIdx.x < ap.world) {
      const uint32_t* flag = fpush      ? p2p::push_flag(ap.peers[ap.rank], (int)threadIdx.x)  // local
                             : one_flag ? reinterpret_cast<const uint32_t*>(ap.peers[threadIdx.x])
                                        : p2p::block_flag(ap.peers[threadIdx.x], ap.seq, blockIdx.x);
      const long long t0 = clock64();
      while ((int32_t)(p2p::ld_acquire_sys(flag) - ap.seq) < 0) {
        if (clock64() - t0 > 20000000000LL) {  // a peer is gone -- report instead of hanging the GPU
          *ap.error_flag = 1;
          timed_out = 1;
          break;
        }
        if (!fpush) __nanosleep(200);
      }
    }
    if (one_flag) {
      if (blockIdx.x == 0) {
        const int lost = __syncthreads_or(timed_out);
        if (threadIdx.x == 0) p2p::st_release_sys(relay, lost ? 0xFFFFFFFFu : ap.seq);  // all ones: a peer is gone
        timed_out = lost;
      } else if (threadIdx.x == 0) {
        uint32_t v;
        while ((int32_t)((v = p2p::ld_acquire_sys(relay)) - ap.seq) < 0 && v != 0xFFFFFFFFu) __nanosleep(100);
        timed_out = (v == 0xFFFFFFFFu) ? 1 : 0;
      }
    }
    peer_lost = __syncthreads_or(timed_out) != 0;
    if (!peer_lost) {
      auto xsum = [&](int e) {  // element e summed over the ranks, in rank order
        if (!fpush) return p2p::peer_sum(ap.peers, ap.world, ap.slot_b, ap.seq, e);
        float v[p2p::PUSH_MAX_RANKS], acc = 0.f;
#pragma unroll
        for (int r = 0; r < p2p::PUSH_MAX_RANKS; ++r)
          v[r] = r < ap.world ? p2p::ld_volatile_f32(p2p::push_recv(ap.peers[ap.rank], ap.slot_b, ap.seq, r) + e) : 0.f;
#pragma unroll
        for (int r = 0; r < p2p::PUSH_MAX_RANKS; ++r)
          if (r < ap.world) acc += v[r];
        return acc;
      };
      if (rg == 0 && p < P && p != log_std_off) t = xsum(p);
      if (blockIdx.x == 0 && rg == 1) {  // the statistics and the log_std entry: written by block 0 on every rank
        if (pl < LRX_PPO_NSTATS) {
          ap.gsum[P + pl] = xsum(P + pl);
        } else if (pl == LRX_PPO_NSTATS && log_std_off >= 0) {
          const float gs = xsum(log_std_off);
          ap.gsum[log_std_off] = gs;
          const float gl = gs + (-ap.h.entropy_loss_coefficient);
          lsg_ss = (double)gl * gl;
        }
      }
    }
  }

  // ---- fused apply (cooperative launch): per-block sum of squares (the log_std entry, owned by
  // block 0, carries the -c_H entropy term exactly as lrx_ppo_apply adds it), grid barrier, every
  // block rebuilds the same global norm in block order and updates ITS 32 parameters.
  double ss = 0;
  if (rg == 0) ss = (double)t * t;
  if (!p2p_on && blockIdx.x == 0 && rg == 4 && pl == 0 && log_std_off >= 0) {
    const float gl = (float)lsg + (-ap.h.entropy_loss_coefficient);
    ss = (double)gl * gl;
  }
  if (p2p_on && rg == 1) ss = lsg_ss;
  ss = warp_sum(ss);
  if (pl == 0) red[rg] = ss;
  __syncthreads();
  if (threadIdx.x == 0) ap.blk_ss[blockIdx.x] = p2p_on ? red[0] + red[1] : red[0] + red[4];
  const int cnt = *ap.count + 1;  // read before the barrier; block 0 stores it after
  cooperative_groups::this_grid().sync();
  double acc = 0;
  for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) acc += __ldcg(ap.blk_ss + b);
  acc = warp_sum(acc);
  __syncthreads();
  if (pl == 0) red[rg] = acc;
  __syncthreads();
  double tt = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) tt += red[i];
  const float norm = (float)sqrt(tt);
  const LrxPpoHyper& h = ap.h;
  // a peer that timed out in ANY block (the flag is sticky) voids the exchange for every block alike
  if (p2p_on) peer_lost = *reinterpret_cast<volatile int*>(ap.error_flag) != 0;
  const float* gsrc = p2p_on ? ap.gsum : gradbuf;  // where the (summed) statistics and log_std entry are
  const float nbad = peer_lost ? 0.f : __ldcg(gsrc + P + 6);
  const bool bad = peer_lost || nbad > 0.f || !isfinite(norm);
  if (!bad && rg == 0 && p < P) {
    const float ls_extra = (log_std_off >= 0) ? -h.entropy_loss_coefficient : 0.f;
    const float bc1 = 1.0f - powf(h.adam_b1, (float)cnt);
    const float bc2 = 1.0f - powf(h.adam_b2, (float)cnt);
    float gp = (p2p_on && p != log_std_off) ? t : __ldcg(gsrc + p) + (p == log_std_off ? ls_extra : 0.f);
    if (!(norm < h.max_grad_norm)) gp = (gp / norm) * h.max_grad_norm;
    const float mn = h.adam_b1 * ap.m[p] + (1.0f - h.adam_b1) * gp;
    const float vn = h.adam_b2 * ap.v[p] + (1.0f - h.adam_b2) * (gp * gp);
    ap.m[p] = mn;
    ap.v[p] = vn;
    const float upd = (mn / bc1) / (sqrtf(vn / bc2) + h.adam_eps);
    ap.params[p] = ap.params[p] + (-h.learning_rate) * upd;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (!bad) *ap.count = cnt;
    else if (ap.nonfinite_flag && !peer_lost) atomicOr(ap.nonfinite_flag, 1);
    if (ap.stats_out && !peer_lost) {
      const float kl = __ldcg(gsrc + P + 0) - 1.0f;
      const float plo = -__ldcg(gsrc + P + 1);
      const float vl = __ldcg(gsrc + P + 2) / 2.0f;
      const float el = -__ldcg(gsrc + P + 3);
      ap.stats_out[0] = kl;
      ap.stats_out[1] = plo + vl * h.value_loss_coefficient + el * h.entropy_loss_coefficient;
      ap.stats_out[2] = plo;
      ap.stats_out[3] = vl;
      ap.stats_out[4] = el;
      ap.stats_out[5] = norm;
      ap.stats_out[6] = nbad;
      ap.stats_out[7] = 0.f;
    }
  }
}

// ===================================================================== clip + Adam
// lrx_ppo_apply: every block rebuilds the global norm of the (all-reduced) gradbuf by itself, in the
// same order (deterministic, no scratch memory, L2-resident reads), then updates its own 256
// parameters.  Cooperative launch: the grid barrier orders block 0's store to *count after every
// block has read it.
__global__ void __launch_bounds__(256)
apply_kernel(float* __restrict__ params, float* __restrict__ m, float* __restrict__ v, int32_t* count,
             const float* __restrict__ g, int P, int log_std_off, LrxPpoHyper h, float* stats_out,
             int32_t* nonfinite_flag, double* blk_ss) {
  __shared__ double red[8];
  const int tid = threadIdx.x;
  // d entropy_loss / d log_std = -1 per row mean -> -c_H, contributed once (not per rank):
  // callers all-reduce gradbuf, so it is added here, after the reduction.
  const float ls_extra = (log_std_off >= 0) ? -h.entropy_loss_coefficient : 0.f;
  const int cnt = *count + 1;
  // global norm: with scratch (lrx_ppo_update) every block sums ITS parameters, the blocks' sums meet at the grid barrier
  // and are added in block order; without (the public lrx_ppo_apply has no workspace) every block walks all of them.
  // Fixed order either way -> deterministic.
  double ss = 0;
  const int first = blk_ss ? blockIdx.x * blockDim.x + tid : tid, stride = blk_ss ? gridDim.x * blockDim.x : blockDim.x;
  for (int p = first; p < P; p += stride) {
    float gp = g[p] + (p == log_std_off ? ls_extra : 0.f);
    ss += (double)gp * gp;
  }
  ss = warp_sum(ss);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  double t = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += red[i];
  if (blk_ss && tid == 0) blk_ss[blockIdx.x] = t;
  cooperative_groups::this_grid().sync();
  if (blk_ss) {
    __shared__ double tot_s;
    double acc = 0;
    for (int b = tid; b < (int)gridDim.x; b += blockDim.x) acc += __ldcg(blk_ss + b);
    acc = warp_sum(acc);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
      double u = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) u += red[i];
      tot_s = u;
    }
    __syncthreads();
    t = tot_s;
  }
  const float norm = (float)sqrt(t);
  const float nbad = g[P + 6];
  const bool bad = nbad > 0.f || !isfinite(norm);
  // the grid is capped at what a cooperative launch can keep resident (launch_apply): walk the parameters with a stride
  if (!bad) {
    const float bc1 = 1.0f - powf(h.adam_b1, (float)cnt);
    const float bc2 = 1.0f - powf(h.adam_b2, (float)cnt);
    for (int p = blockIdx.x * blockDim.x + tid; p < P; p += gridDim.x * blockDim.x) {
      float gp = g[p] + (p == log_std_off ? ls_extra : 0.f);
      if (!(norm < h.max_grad_norm)) gp = (gp / norm) * h.max_grad_norm;
      const float mn = h.adam_b1 * m[p] + (1.0f - h.adam_b1) * gp;
      const float vn = h.adam_b2 * v[p] + (1.0f - h.adam_b2) * (gp * gp);
      m[p] = mn;
      v[p] = vn;
      const float upd = (mn / bc1) / (sqrtf(vn / bc2) + h.adam_eps);
      params[p] = params[p] + (-h.learning_rate) * upd;
    }
  }
  if (blockIdx.x == 0 && tid == 0) {
    if (!bad) *count = cnt;
    else if (nonfinite_flag) atomicOr(nonfinite_flag, 1);
    if (stats_out) {
      const float kl = g[P + 0] - 1.0f;
      const float pl = -g[P + 1];
      const float vl = g[P + 2] / 2.0f;
      const float el = -g[P + 3];
      stats_out[0] = kl;
      stats_out[1] = pl + vl * h.value_loss_coefficient + el * h.entropy_loss_coefficient;
      stats_out[2] = pl;
      stats_out[3] = vl;
      stats_out[4] = el;
      stats_out[5] = norm;
      stats_out[6] = nbad;
      stats_out[7] = 0.f;
    }
  }
}

// The same with the exchange inside: every block first sums ITS 256 gradient elements over the ranks'
// exchange slots in NVLink peer memory (lrx_p2p.cuh; rank order, bit-identical on every rank), the blocks' sums of
// squares meet at the grid barrier, then clip + Adam on the block's own elements -- gradient all-reduce, global norm
// and optimiser step in one cooperative launch, no collective library call.  gsum[P + 8] receives the summed
// gradbuf (the stat sums are read from it); blk_ss holds one double per block.
__global__ void __launch_bounds__(256)
apply_p2p_kernel(float* __restrict__ params, float* __restrict__ m, float* __restrict__ v, int32_t* count,
                 char* const* __restrict__ peers, int rank, int world, size_t slot_b, uint32_t seq, int* error_flag,
                 float* __restrict__ gsum, double* __restrict__ blk_ss, int P, int log_std_off, LrxPpoHyper h,
                 float* stats_out, int32_t* nonfinite_flag, int push) {
  __shared__ double red[8];
  __shared__ float norm_s;
  asm volatile("griddepcontrol.wait;" ::: "memory");  // programmatic dependent launch: the own slot is the previous kernel's output
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int tid = threadIdx.x;
  bool arrived;
  if (push) {  // the peers wrote their gradients and arrival flags into THIS rank's buffer: local polling, local loads
    int timed_out = 0;
    if (tid < world) {
      const uint32_t* flag = p2p::push_flag(peers[rank], tid);
      const long long t0 = clock64();
      while ((int32_t)(p2p::ld_acquire_sys(flag) - seq) < 0) {
        if (clock64() - t0 > 20000000000LL) {
          *error_flag = 1;
          timed_out = 1;
          break;
        }
      }
    }
    arrived = __syncthreads_or(timed_out) == 0;
  } else {
    arrived = p2p::publish_and_wait(peers, rank, world, seq, blockIdx.x == 0, error_flag);
  }
  const float ls_extra = (log_std_off >= 0) ? -h.entropy_loss_coefficient : 0.f;
  const int cnt = *count + 1;
  // pass 1 (grid stride; the grid is capped at what a cooperative launch keeps resident): rank-ordered sum of every
  // element over the peers' slots -> gsum, this block's share of the sum of squares
  double ss = 0;
  if (arrived) {  // a lost peer's slot may be stale or half written: never summed
    for (int p = blockIdx.x * blockDim.x + tid; p < P + LRX_PPO_NSTATS; p += gridDim.x * blockDim.x) {
      float sum;
      if (push) {
        sum = 0.f;
        for (int r = 0; r < world; ++r) sum += p2p::ld_volatile_f32(p2p::push_recv(peers[rank], slot_b, seq, r) + p);  // rank order
      } else {
        sum = p2p::peer_sum(peers, world, slot_b, seq, p);
      }
      gsum[p] = sum;
      if (p < P) {
        const float gp = sum + (p == log_std_off ? ls_extra : 0.f);
        ss += (double)gp * gp;
      }
    }
  }
  ss = warp_sum(ss);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  if (tid == 0) {
    double t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    blk_ss[blockIdx.x] = t;
  }
  cooperative_groups::this_grid().sync();
  if (tid == 0) {
    double t = 0;
    for (unsigned b = 0; b < gridDim.x; ++b) t += blk_ss[b];  // block order: deterministic (same grid on every rank)
    norm_s = (float)sqrt(t);
  }
  __syncthreads();
  const float norm = norm_s;
  // a peer that timed out in ANY block (the flag is sticky) voids the exchange for every block alike: no parameter, moment
  // or count update and no statistics; the host raises when it reads the flag (GradientExchange.check)
  const bool peer_lost = *reinterpret_cast<volatile int*>(error_flag) != 0;
  const float nbad = peer_lost ? 0.f : gsum[P + 6];
  const bool bad = peer_lost || nbad > 0.f || !isfinite(norm);
  if (!bad) {  // pass 2: clip + Adam on the block's own elements (written by this very thread in pass 1)
    const float bc1 = 1.0f - powf(h.adam_b1, (float)cnt);
    const float bc2 = 1.0f - powf(h.adam_b2, (float)cnt);
    for (int p = blockIdx.x * blockDim.x + tid; p < P; p += gridDim.x * blockDim.x) {
      float gp = gsum[p] + (p == log_std_off ? ls_extra : 0.f);
      if (!(norm < h.max_grad_norm)) gp = (gp / norm) * h.max_grad_norm;
      const float mn = h.adam_b1 * m[p] + (1.0f - h.adam_b1) * gp;
      const float vn = h.adam_b2 * v[p] + (1.0f - h.adam_b2) * (gp * gp);
      m[p] = mn;
      v[p] = vn;
      const float upd = (mn / bc1) / (sqrtf(vn / bc2) + h.adam_eps);
      params[p] = params[p] + (-h.learning_rate) * upd;
    }
  }
  if (blockIdx.x == 0 && tid == 0) {
    if (!bad) *count = cnt;
    else if (nonfinite_flag && !peer_lost) atomicOr(nonfinite_flag, 1);
    if (stats_out && !peer_lost) {
      const float kl = gsum[P + 0] - 1.0f;
      const float pl = -gsum[P + 1];
      const float vl = gsum[P + 2] / 2.0f;
      const float el = -gsum[P + 3];
      stats_out[0] = kl;
      stats_out[1] = pl + vl * h.value_loss_coefficient + el * h.entropy_loss_coefficient;
      stats_out[2] = pl;
      stats_out[3] = vl;
      stats_out[4] = el;
      stats_out[5] = norm;
      stats_out[6] = nbad;
      stats_out[7] = 0.f;
    }
  }
}

static int launch_apply(float* params, float* m, float* v, int32_t* count, const float* g, int P, int log_std_off,
                        const LrxPpoHyper& h, float* stats_out, int32_t* nonfinite_flag, cudaStream_t s,
                        double* blk_ss = nullptr) {
  LrxPpoHyper hh = h;
  void* args[] = {(void*)&params, (void*)&m, (void*)&v, (void*)&count, (void*)&g, (void*)&P, (void*)&log_std_off,
                  (void*)&hh, (void*)&stats_out, (void*)&nonfinite_flag, (void*)&blk_ss};
  // every block must be resident at once (grid barrier): at most 4 blocks of 256 threads per SM, parameters walked with a
  // grid stride (a 512-wide policy has > 300 k parameters, more blocks than a cooperative launch accepts)
  // without scratch every block reads the whole gradient for the norm: keep that to one block per SM
  const int cap = (blk_ss ? 4 : 1) * lrx_sm_count(), want = (P + 255) / 256;
  LRX_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)apply_kernel, dim3(want < cap ? want : cap), dim3(256), args, 0, s));
  return LRX_OK;
}

// ========================================================================= workspace
struct Workspace {
  int64_t rows;      // row capacity of one chunk (multiple of 4)
  int splits;
  int loss_blocks;
  size_t P_pad;
  float* obsT;
  float* act_out[3][LRX_MAX_LAYERS];  // post-activation outputs (+ ones row), [out+1][rows]
  float* dA; float* dB; float* dfeat; float* dval; float* dhd;
  float* act_f; int32_t* act_i; float* logp; float* adv; float* ret; float* val;
  float* partial;        // [splits][P_pad]
  double* loss_partial;  // [loss_blocks][8]
  size_t total;
};
constexpr int LRX_MAX_BATCHES = 4096;

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static Workspace carve(const LrxPolicyDesc& pol, int64_t B, void* base) {
  Workspace w{};
  static const int64_t chunk_rows = [] { const char* e = getenv("LRX_CHUNK_ROWS"); return e ? atoll(e) : LRX_CHUNK_ROWS; }();
  static const int64_t split_rows = [] { const char* e = getenv("LRX_SPLIT_ROWS"); return e ? atoll(e) : 1024LL; }();
  int64_t rows = B < chunk_rows ? B : chunk_rows;
  rows = (rows + 3) / 4 * 4;
  if (rows < 4) rows = 4;
  w.rows = rows;
  int sp = (int)((rows + split_rows - 1) / split_rows);
  w.splits = sp < 1 ? 1 : (sp > LRX_MAX_SPLIT ? LRX_MAX_SPLIT : sp);
  w.loss_blocks = (int)((rows + LRX_LOSS_THREADS - 1) / LRX_LOSS_THREADS);
  int np = 0;
  lrx_policy_layers(pol, nullptr, nullptr, &np);
  w.P_pad = align_up((size_t)np + LRX_PPO_NSTATS, 32);
  size_t off = 0;
  char* b = (char*)base;
  auto take = [&](size_t bytes) { void* p = b ? b + off : nullptr; off += align_up(bytes, 256); return p; };
  int maxw = 1;
  for (int c = 0; c < 3; ++c)
    for (int i = 0; i <= pol.n_layers[c]; ++i) maxw = pol.dims[c][i] > maxw ? pol.dims[c][i] : maxw;
  w.obsT = (float*)take((size_t)(pol.obs_dim + 1) * rows * 4);
  for (int c = 0; c < 3; ++c)
    for (int i = 0; i < pol.n_layers[c]; ++i) w.act_out[c][i] = (float*)take((size_t)(pol.dims[c][i + 1] + 1) * rows * 4);
  w.dA = (float*)take((size_t)maxw * rows * 4);
  w.dB = (float*)take((size_t)maxw * rows * 4);
  w.dfeat = (float*)take((size_t)pol.dims[0][pol.n_layers[0]] * rows * 4);
  w.dval = (float*)take((size_t)rows * 4);
  w.dhd = (float*)take((size_t)pol.n_out * rows * 4);
  w.act_f = (float*)take(rows * 4);
  w.act_i = (int32_t*)w.act_f;
  w.logp = (float*)take(rows * 4);
  w.adv = (float*)take(rows * 4);
  w.ret = (float*)take(rows * 4);
  w.val = (float*)take(rows * 4);
  w.partial = (float*)take((size_t)w.splits * w.P_pad * 4);
  w.loss_partial = (double*)take((size_t)w.loss_blocks * 8 * sizeof(double) * ((B + rows - 1) / rows));
  w.total = off;
  return w;
}

size_t lrx_fused_workspace_bytes(const LrxPolicyDesc* pol, int64_t B);

// Workspace = [max(generic scratch, fused scratch)] [adv_sums of lrx_ppo_update] [gradbuf of lrx_ppo_update].
struct WorkspaceTail {
  size_t adv_sums_off, gradbuf_off, scratch_off, total;
};
constexpr size_t LRX_APPLY_SCRATCH = 8192;  // per-block sums of squares (<= 1000 blocks) + ticket
static WorkspaceTail workspace_tail(const LrxPolicyDesc* pol, int64_t B) {
  Workspace w = carve(*pol, B, nullptr);
  size_t a = w.total;
  size_t f = lrx_fused_workspace_bytes(pol, B);
  WorkspaceTail t{};
  t.adv_sums_off = a > f ? a : f;
  t.gradbuf_off = t.adv_sums_off + align_up((size_t)LRX_MAX_BATCHES * 2 * sizeof(double), 256);
  t.scratch_off = t.gradbuf_off + align_up(w.P_pad * 4, 256);
  t.total = t.scratch_off + LRX_APPLY_SCRATCH;
  return t;
}

extern "C" size_t lrx_ppo_workspace_bytes(const LrxPolicyDesc* pol, int64_t B) {
  if (!pol || lrx_validate_policy(pol) != LRX_OK || B <= 0) return 0;
  return workspace_tail(pol, B).total;
}

int lrx_launch_reduce_partials(const float* partial, int splits, long long stride, int P, const double* loss_partial,
                               int loss_blocks, int log_std_off, float ent_coef, double B_global, float* gradbuf,
                               const LrxApplyArgs* apply, cudaStream_t s) {
  LrxApplyArgs ap{};
  const int grid = (P + 31) / 32;
  if (apply && !apply->params) {  // no update, only the early publish of the slot
    ap = *apply;
    apply = nullptr;
  }
  if (apply && grid <= 4 * lrx_sm_count()) {
    // fused reduce + clip + Adam: needs every block resident at once (grid barrier) -> cooperative launch
    ap = *apply;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(256);
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = lrx_pdl_enabled() ? 2 : 1;
    LRX_CUDA_CHECK(cudaLaunchKernelEx(&cfg, reduce_partials_kernel, partial, splits, stride, P, loss_partial, loss_blocks,
                                      log_std_off, ent_coef, B_global, gradbuf, ap));
    return LRX_OK;
  }
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(256);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = lrx_pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    LRX_CUDA_CHECK(cudaLaunchKernelEx(&cfg, reduce_partials_kernel, partial, splits, stride, P, loss_partial, loss_blocks,
                                      log_std_off, ent_coef, B_global, gradbuf, ap));
  }
  if (apply)  // too many blocks for one resident wave of the fused kernel: separate apply launch
    return launch_apply(apply->params, apply->m, apply->v, apply->count, gradbuf, P, log_std_off, apply->h, apply->stats_out,
                        apply->nonfinite_flag, s, apply->blk_ss);
  return LRX_OK;
}

// ======================================================================== host driver
extern "C" int lrx_ppo_adv_sums(const float* adv, int32_t N, int32_t T, const int32_t* idx,
                                int32_t num_batches, int64_t B, double* sums, lrx_stream_t stream) {
  LRX_REQUIRE(adv && idx && sums && N > 0 && T > 0 && num_batches >= 0 && B > 0, LRX_ERR_INVALID_ARG,
              "lrx_ppo_adv_sums: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (num_batches == 0) return LRX_OK;
  unsigned S = 8;  // CTAs per minibatch: enough to fill the SMs, never more than the rows can feed
  while (S > 1 && ((long long)num_batches * S > 2LL * lrx_sm_count() || B < (long long)S * 8 * 1024)) S >>= 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)num_batches * S);
  cfg.blockDim = dim3(1024);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LRX_CUDA_CHECK(cudaLaunchKernelEx(&cfg, adv_sums_kernel, adv, (int)N, (int)T, idx, (long long)B, sums));
  return LRX_OK;
}

static int validate_update_args(const LrxPolicyDesc* pol, const LrxBatchView* buf, const LrxPpoHyper* h) {
  int rc = lrx_validate_policy(pol);
  if (rc) return rc;
  LRX_REQUIRE(buf && h, LRX_ERR_INVALID_ARG, "NULL LrxBatchView / LrxPpoHyper");
  LRX_REQUIRE(buf->obs && buf->act && buf->logp && buf->val && buf->ret && buf->adv && buf->N > 0 && buf->T > 0,
              LRX_ERR_INVALID_ARG, "LrxBatchView has NULL planes or empty shape");
  LRX_REQUIRE((int64_t)buf->N * buf->T < ((int64_t)1 << 31), LRX_ERR_UNSUPPORTED,
              "N*T must fit int32 (reference indices are int32)");
  return LRX_OK;
}

int lrx_ppo_grad_fused(const LrxPolicyDesc* pol, const float* params, const LrxBatchView* buf,
                       const int32_t* idx, int64_t B, int64_t B_global, const double* adv_sums,
                       const LrxPpoHyper* h, float* gradbuf, void* workspace, size_t workspace_bytes,
                       cudaStream_t stream, bool* handled, const LrxApplyArgs* apply);
int lrx_ppo_grad_tc128(const LrxPolicyDesc* pol, const float* params, const LrxBatchView* buf, const int32_t* idx,
                       int64_t B, int64_t B_global, const double* adv_sums, const LrxPpoHyper* h, float* gradbuf,
                       void* workspace, size_t workspace_bytes, cudaStream_t stream, bool* handled,
                       const LrxApplyArgs* apply);
int lrx_ppo_grad_pp64(const LrxPolicyDesc* pol, const float* params, const LrxBatchView* buf, const int32_t* idx,
                      int64_t B, int64_t B_global, const double* adv_sums, const LrxPpoHyper* h, float* gradbuf,
                      void* workspace, size_t workspace_bytes, cudaStream_t stream, bool* handled,
                      const LrxApplyArgs* apply);
int lrx_ppo_grad_tc128c(const LrxPolicyDesc* pol, const float* params, const LrxBatchView* buf, const int32_t* idx,
                        int64_t B, int64_t B_global, const double* adv_sums, const LrxPpoHyper* h, float* gradbuf,
                        void* workspace, size_t workspace_bytes, cudaStream_t stream, bool* handled,
                        const LrxApplyArgs* apply);
// Default-width policies, by lrx_fused_enabled(): 4 = 128-row kernel with composite weights (lrx_update_tc128c.cu, the default),
// 3 = two half tiles in flight (lrx_update_pp64.cu), 2 = 128-row kernel (lrx_update_tc128.cu), 1 = first generation.
// Default-width policies: third-generation (two half tiles in flight, lrx_update_pp64.cu), second-generation tensor-core kernel (lrx_update_tc128.cu), else the first-generation one
// (lrx_update_fused.cu); `handled` stays false for other widths (or with the kernels switched off).
static int lrx_ppo_grad_tensor(const LrxPolicyDesc* pol, const float* params, const LrxBatchView* buf, const int32_t* idx,
                               int64_t B, int64_t B_global, const double* adv_sums, const LrxPpoHyper* h, float* gradbuf,
                               void* workspace, size_t workspace_bytes, cudaStream_t stream, bool* handled,
                               const LrxApplyArgs* apply) {
  int rc = lrx_ppo_grad_tc128c(pol, params, buf, idx, B, B_global, adv_sums, h, gradbuf, workspace, workspace_bytes, stream,
                               handled, apply);
  if (rc || *handled) return rc;
  rc = lrx_ppo_grad_pp64(pol, params, buf, idx, B, B_global, adv_sums, h, gradbuf, workspace, workspace_bytes, stream,
                         handled, apply);
  if (rc || *handled) return rc;
  rc = lrx_ppo_grad_tc128(pol, params, buf, idx, B, B_global, adv_sums, h, gradbuf, workspace, workspace_bytes, stream,
                          handled, apply);
  if (rc || *handled) return rc;
  return lrx_ppo_grad_fused(pol, params, buf, idx, B, B_global, adv_sums, h, gradbuf, workspace, workspace_bytes, stream,
                            handled, apply);
}

extern "C" int lrx_ppo_grad(const LrxPolicyDesc* pol, const float* params, const LrxBatchView* buf,
                            const int32_t* idx, int64_t B, int64_t B_global, const double* adv_sums,
                            const LrxPpoHyper* h, float* gradbuf, void* workspace,
                            size_t workspace_bytes, lrx_stream_t stream) {
  int rc = validate_update_args(pol, buf, h);
  if (rc) return rc;
  LRX_REQUIRE(params && idx && gradbuf && workspace && B > 0 && B_global >= B, LRX_ERR_INVALID_ARG,
              "lrx_ppo_grad: bad argument");
  LRX_REQUIRE(!h->normalize_advantages || adv_sums, LRX_ERR_INVALID_ARG,
              "lrx_ppo_grad: adv_sums required when normalize_advantages is set");
  LRX_REQUIRE(workspace_bytes >= workspace_tail(pol, B).adv_sums_off, LRX_ERR_INVALID_ARG,
              "lrx_ppo_grad: workspace too small (%zu < %zu)", workspace_bytes, workspace_tail(pol, B).adv_sums_off);
  cudaStream_t s = (cudaStream_t)stream;

  bool handled = false;
  rc = lrx_ppo_grad_tensor(pol, params, buf, idx, B, B_global, adv_sums, h, gradbuf, workspace, workspace_bytes, s, &handled, nullptr);
  if (rc) return rc;
  if (handled) return LRX_OK;

  Workspace w = carve(*pol, B, workspace);
  PolicyTable tab = lrx_make_table(*pol);
  const int P = pol->n_params;
  const int ld = (int)w.rows;
  const int nchunks = (int)((B + w.rows - 1) / w.rows);

  auto input_of = [&](int c, int i) -> float* {
    if (i > 0) return w.act_out[c][i - 1];
    return c == 0 ? w.obsT : w.act_out[0][pol->n_layers[0] - 1];
  };

  for (int ch = 0; ch < nchunks; ++ch) {
    const int64_t r0 = (int64_t)ch * w.rows;
    const int rows = (int)((B - r0) < w.rows ? (B - r0) : w.rows);
    const int splits = w.splits;
    const int k_chunk = ((rows + splits - 1) / splits + 15) / 16 * 16;
    gather_kernel<<<(rows + 255) / 256, 256, 0, s>>>(*buf, idx + r0, rows, ld, pol->obs_dim, pol->is_box, w.obsT,
                                                     w.act_f, w.act_i, w.logp, w.adv, w.ret, w.val);
    LRX_LAUNCH_CHECK();

    // ---- forward: Out[out x rows] = act(W[out x in] * In[in x rows] + b), plus a ones row
    for (int c = 0; c < 3; ++c) {
      for (int i = 0; i < tab.count[c]; ++i) {
        const LrxLayer& L = tab.L[tab.first[c] + i];
        GemmArgs g{};
        g.A = params + L.w_off; g.lda = L.in;
        g.B = input_of(c, i); g.ldb = ld;
        g.C = w.act_out[c][i]; g.ldc = ld;
        g.M = L.out; g.N = rows; g.K = L.in;
        g.bias = params + L.b_off; g.relu = L.relu; g.ones_row = 1;
        g.k_chunk = L.in;
        if ((rc = launch_gemm(g, false, false, 1, s))) return rc;
      }
    }

    // ---- loss: dL/dvalue -> dval [1][rows], dL/dhead -> dhd [n_out][rows]
    LossArgs la{};
    la.value = w.act_out[1][pol->n_layers[1] - 1];
    la.head = w.act_out[2][pol->n_layers[2] - 1];
    la.ld = ld;
    la.act_f = w.act_f; la.act_i = w.act_i;
    la.logp_old = w.logp; la.adv = w.adv; la.ret = w.ret; la.val_old = w.val;
    la.dvalue = w.dval; la.dhead = w.dhd;
    la.rows = rows; la.n_out = pol->n_out; la.is_box = pol->is_box;
    la.log_std_ptr = pol->is_box ? params + tab.log_std_off : nullptr;
    la.adv_sums = adv_sums; la.B_global = (double)B_global; la.h = *h;
    la.partial = w.loss_partial + (size_t)ch * w.loss_blocks * 8;
    const int lblocks = (rows + LRX_LOSS_THREADS - 1) / LRX_LOSS_THREADS;
    loss_kernel<<<lblocks, LRX_LOSS_THREADS, 0, s>>>(la);
    LRX_LAUNCH_CHECK();

    // ---- backward through one chain, starting from dOut = d_cur [out_last x rows]
    auto backprop_chain = [&](int c, float* d_cur, float* d_final, bool accumulate_final) -> int {
      for (int i = tab.count[c] - 1; i >= 0; --i) {
        const LrxLayer& L = tab.L[tab.first[c] + i];
        float* in = input_of(c, i);
        // [dW | db] = dOut[out x rows] * [In ; 1]^T, split-K over the rows
        GemmArgs g{};
        g.A = d_cur; g.lda = ld;
        g.B = in; g.ldb = ld;
        g.C = w.partial + L.w_off; g.ldc = L.in;
        g.C2 = w.partial + L.b_off; g.n_main = L.in;
        g.M = L.out; g.N = L.in + 1; g.K = rows;
        g.k_chunk = k_chunk; g.c_split_stride = (long long)w.P_pad;
        g.accumulate = ch > 0;
        int rcl = launch_gemm(g, false, true, splits, s);
        if (rcl) return rcl;
        if (i == 0 && c == 0) break;  // no gradient w.r.t. the observations
        // dIn[in x rows] = W^T * dOut, masked by the ReLU of the layer that produced In
        GemmArgs d{};
        d.A = params + L.w_off; d.lda = L.in;
        d.B = d_cur; d.ldb = ld;
        d.M = L.in; d.N = rows; d.K = L.out; d.k_chunk = L.out;
        float* dst;
        if (i == 0) {
          dst = d_final;  // gradient w.r.t. the encoder output (identity activation)
          d.accumulate = accumulate_final;
        } else {
          dst = (d_cur == w.dA) ? w.dB : w.dA;
          if (tab.L[tab.first[c] + i - 1].relu) { d.mask = in; d.ldm = ld; }
        }
        d.C = dst; d.ldc = ld;
        rcl = launch_gemm(d, true, false, 1, s);
        if (rcl) return rcl;
        d_cur = dst;
      }
      return LRX_OK;
    };
    if ((rc = backprop_chain(1, w.dval, w.dfeat, false))) return rc;
    if ((rc = backprop_chain(2, w.dhd, w.dfeat, true))) return rc;
    if ((rc = backprop_chain(0, w.dfeat, nullptr, false))) return rc;
  }

  reduce_partials_kernel<<<(P + 31) / 32, 256, 0, s>>>(w.partial, w.splits, (long long)w.P_pad, P, w.loss_partial,
                                                        nchunks * w.loss_blocks, tab.log_std_off,
                                                        h->entropy_loss_coefficient, (double)B_global, gradbuf,
                                                        LrxApplyArgs{});
  LRX_LAUNCH_CHECK();
  return LRX_OK;
}

extern "C" int lrx_ppo_apply(const LrxPolicyDesc* pol, float* params, float* adam_m, float* adam_v,
                             int32_t* adam_count, const float* gradbuf, const LrxPpoHyper* h,
                             float* stats_out, int32_t* nonfinite_flag, lrx_stream_t stream) {
  int rc = lrx_validate_policy(pol);
  if (rc) return rc;
  LRX_REQUIRE(params && adam_m && adam_v && adam_count && gradbuf && h, LRX_ERR_INVALID_ARG,
              "lrx_ppo_apply: NULL pointer");
  PolicyTable tab = lrx_make_table(*pol);
  int rc2 = launch_apply(params, adam_m, adam_v, adam_count, gradbuf, pol->n_params, tab.log_std_off, *h, stats_out,
                         nonfinite_flag, (cudaStream_t)stream);
  if (rc2) return rc2;
  LRX_LAUNCH_CHECK();
  return LRX_OK;
}

extern "C" size_t lrx_ppo_apply_p2p_scratch_bytes(const LrxPolicyDesc* pol) {
  if (!pol) return 0;
  const size_t n = (size_t)pol->n_params + LRX_PPO_NSTATS;
  return align_up(n * 4, 256) + ((n + 255) / 256) * sizeof(double);
}

static int apply_p2p_impl(const LrxPolicyDesc* pol, float* params, float* adam_m, float* adam_v, int32_t* adam_count,
                          void* const* peer_buffers, int32_t rank, int32_t world, uint32_t seq, const LrxPpoHyper* h,
                          float* stats_out, int32_t* nonfinite_flag, int32_t* error_flag, void* scratch, size_t scratch_bytes,
                          lrx_stream_t stream, int push);
extern "C" int lrx_ppo_apply_p2p(const LrxPolicyDesc* pol, float* params, float* adam_m, float* adam_v,
                                 int32_t* adam_count, void* const* peer_buffers, int32_t rank, int32_t world,
                                 uint32_t seq, const LrxPpoHyper* h, float* stats_out, int32_t* nonfinite_flag,
                                 int32_t* error_flag, void* scratch, size_t scratch_bytes, lrx_stream_t stream) {
  return apply_p2p_impl(pol, params, adam_m, adam_v, adam_count, peer_buffers, rank, world, seq, h, stats_out, nonfinite_flag,
                        error_flag, scratch, scratch_bytes, stream, 0);
}
static int apply_p2p_impl(const LrxPolicyDesc* pol, float* params, float* adam_m, float* adam_v, int32_t* adam_count,
                          void* const* peer_buffers, int32_t rank, int32_t world, uint32_t seq, const LrxPpoHyper* h,
                          float* stats_out, int32_t* nonfinite_flag, int32_t* error_flag, void* scratch, size_t scratch_bytes,
                          lrx_stream_t stream, int push) {
  int rc = lrx_validate_policy(pol);
  if (rc) return rc;
  LRX_REQUIRE(params && adam_m && adam_v && adam_count && peer_buffers && h && error_flag && scratch,
              LRX_ERR_INVALID_ARG, "lrx_ppo_apply_p2p: NULL pointer");
  LRX_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world && seq != 0, LRX_ERR_INVALID_ARG,
              "lrx_ppo_apply_p2p: bad rank / world / seq");
  LRX_REQUIRE(scratch_bytes >= lrx_ppo_apply_p2p_scratch_bytes(pol), LRX_ERR_INVALID_ARG,
              "lrx_ppo_apply_p2p: scratch too small");
  PolicyTable tab = lrx_make_table(*pol);
  int P = pol->n_params, n = P + LRX_PPO_NSTATS, log_std_off = tab.log_std_off;
  char* const* peers = reinterpret_cast<char* const*>(peer_buffers);
  size_t slot_b = p2p::slot_bytes(n);
  float* gsum = (float*)scratch;
  double* blk_ss = (double*)((char*)scratch + align_up((size_t)n * 4, 256));
  LrxPpoHyper hh = *h;
  int rk = rank, wd = world;
  uint32_t sq = seq;
  void* args[] = {(void*)&params, (void*)&adam_m, (void*)&adam_v, (void*)&adam_count, (void*)&peers, (void*)&rk,
                  (void*)&wd, (void*)&slot_b, (void*)&sq, (void*)&error_flag, (void*)&gsum, (void*)&blk_ss, (void*)&P,
                  (void*)&log_std_off, (void*)&hh, (void*)&stats_out, (void*)&nonfinite_flag, (void*)&push};
  const int cap = 4 * lrx_sm_count(), want = (n + 255) / 256;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(want < cap ? want : cap));
  cfg.blockDim = dim3(256);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = lrx_pdl_enabled() ? 2 : 1;
  LRX_CUDA_CHECK(cudaLaunchKernelExC(&cfg, (const void*)apply_p2p_kernel, args));
  LRX_LAUNCH_CHECK();
  return LRX_OK;
}

// lrx_ppo_grad for the NVLink exchange: the gradient goes into slot `seq & 1` of this rank's exchange buffer and the reduction
// kernel's last block publishes `seq` in the buffer's flag, so the peers can read the slot while this rank's
// lrx_ppo_apply_p2p (which publishes the same value again: harmless) is still being launched.
static int grad_p2p_impl(const LrxPolicyDesc* pol, const float* params, const LrxBatchView* buf, const int32_t* idx, int64_t B,
                         int64_t B_global, const double* adv_sums, const LrxPpoHyper* h, void* own_buffer, uint32_t seq,
                         void* workspace, size_t workspace_bytes, lrx_stream_t stream, void* const* push_peers, int rank,
                         int world, bool* pushed);
extern "C" int lrx_ppo_grad_p2p(const LrxPolicyDesc* pol, const float* params, const LrxBatchView* buf, const int32_t* idx,
                                int64_t B, int64_t B_global, const double* adv_sums, const LrxPpoHyper* h, void* own_buffer,
                                uint32_t seq, void* workspace, size_t workspace_bytes, lrx_stream_t stream) {
  bool pushed = false;
  return grad_p2p_impl(pol, params, buf, idx, B, B_global, adv_sums, h, own_buffer, seq, workspace, workspace_bytes, stream,
                       nullptr, 0, 1, &pushed);
}
// push_peers != nullptr: the reduction also writes the gradient into every rank's receive area and sets the arrival flags
// (*pushed = true); policies the fused gradient kernels do not cover fall back to the plain slot (*pushed = false).
static int grad_p2p_impl(const LrxPolicyDesc* pol, const float* params, const LrxBatchView* buf, const int32_t* idx, int64_t B,
                         int64_t B_global, const double* adv_sums, const LrxPpoHyper* h, void* own_buffer, uint32_t seq,
                         void* workspace, size_t workspace_bytes, lrx_stream_t stream, void* const* push_peers, int rank,
                         int world, bool* pushed) {
  *pushed = false;
  int rc = validate_update_args(pol, buf, h);
  if (rc) return rc;
  LRX_REQUIRE(params && idx && own_buffer && workspace && B > 0 && B_global >= B && seq != 0, LRX_ERR_INVALID_ARG,
              "lrx_ppo_grad_p2p: bad argument");
  LRX_REQUIRE(workspace_bytes >= workspace_tail(pol, B).adv_sums_off, LRX_ERR_INVALID_ARG,
              "lrx_ppo_grad_p2p: workspace too small");
  const int n = pol->n_params + LRX_PPO_NSTATS;
  char* own = static_cast<char*>(own_buffer);
  float* slot = reinterpret_cast<float*>(own + p2p::HEADER + (size_t)(seq & 1u) * p2p::slot_bytes(n));
  LrxApplyArgs ap{};
  ap.publish_flag = reinterpret_cast<uint32_t*>(own);
  ap.publish_seq = seq;
  ap.ticket = reinterpret_cast<unsigned*>(own + 64);  // inside the global flag's line; zeroed by lrx_p2p_alloc, self-resetting
  if (push_peers) {
    ap.push = 1;
    ap.peers = reinterpret_cast<char* const*>(push_peers);
    ap.rank = rank;
    ap.world = world;
    ap.slot_b = p2p::slot_bytes(n);
  }
  bool handled = false;
  rc = lrx_ppo_grad_tensor(pol, params, buf, idx, B, B_global, adv_sums, h, slot, workspace, workspace_bytes,
                           (cudaStream_t)stream, &handled, &ap);
  if (rc) return rc;
  if (handled) {
    *pushed = push_peers != nullptr;
    return LRX_OK;
  }
  return lrx_ppo_grad(pol, params, buf, idx, B, B_global, adv_sums, h, slot, workspace, workspace_bytes, stream);
}

// One epoch of the multi-GPU update driven from C: per minibatch lrx_ppo_grad_p2p + lrx_ppo_apply_p2p with the sequence
// numbers seq64 + 1 .. seq64 + num_batches (folded into 1 .. 2^32 - 2 with an even period, so consecutive exchanges
// always alternate between the two slots).  `adv_sums` [num_batches][2] are the (already all-reduced) advantage sums.
// Per-minibatch host work is what bounds a data-parallel run once the kernels are fast: 2 x 320 interpreter round trips
// per iteration cost more than the exchange itself.
extern "C" int lrx_ppo_grad_apply_p2p(const LrxPolicyDesc* pol, float* params, float* adam_m, float* adam_v,
                                      int32_t* adam_count, const LrxBatchView* buf, const int32_t* idx, int64_t B,
                                      int64_t B_global, const double* adv_sums, const LrxPpoHyper* h,
                                      void* const* peer_buffers, void* own_buffer, int32_t rank, int32_t world,
                                      uint32_t seq, float* stats_out, int32_t* nonfinite_flag, int32_t* error_flag,
                                      void* workspace, size_t workspace_bytes, lrx_stream_t stream, int32_t* handled);
extern "C" int lrx_ppo_update_p2p(const LrxPolicyDesc* pol, float* params, float* adam_m, float* adam_v, int32_t* adam_count,
                                  const LrxBatchView* buf, const int32_t* idx, int32_t num_batches, int64_t B,
                                  int64_t B_global, const double* adv_sums, const LrxPpoHyper* h,
                                  const float* learning_rates_host, void* const* peer_buffers, void* own_buffer,
                                  int32_t rank, int32_t world, uint64_t seq64, float* stats_out, int32_t* nonfinite_flag,
                                  int32_t* error_flag, void* scratch, size_t scratch_bytes, void* workspace,
                                  size_t workspace_bytes, lrx_stream_t stream) {
  LRX_REQUIRE(num_batches >= 0 && num_batches <= LRX_MAX_BATCHES && idx && stats_out && h, LRX_ERR_INVALID_ARG,
              "lrx_ppo_update_p2p: bad argument");
  LrxPpoHyper hh = *h;
  // LRX_P2P_PUSH=1: push exchange (p2p::PUSH_MAX_RANKS receive areas per slot).  Measured equal to the pull protocol at 2
  // GPUs (197.7 vs 197.3 - 200 us per minibatch), so the simpler pull protocol stays the default.
  int push_env = 0;
  {
    const char* e = getenv("LRX_P2P_PUSH");  // read per call (once per epoch): the tests switch variants within a process
    push_env = (e && e[0] == '1') ? 1 : 0;
  }
  const bool use_push = push_env && world <= p2p::PUSH_MAX_RANKS;
  // default: TWO launches per minibatch (lrx_ppo_grad_apply_p2p; 191 us against 193 us for the three calls below at 2 GPUs);
  // LRX_P2P_FUSED=0 or a policy the fused gradient kernels do not cover: three
  bool fused = !use_push;
  {
    const char* e = getenv("LRX_P2P_FUSED");
    if (e && e[0] == '0') fused = false;
  }
  for (int b = 0; b < num_batches; ++b) {
    const uint32_t seq = (uint32_t)((seq64 + (uint64_t)b) % 0xFFFFFFFEull) + 1u;  // = ((seq64 + b + 1) - 1) % (2^32 - 2) + 1
    if (learning_rates_host) hh.learning_rate = learning_rates_host[b];
    if (fused) {
      int32_t handled = 0;
      const int rc = lrx_ppo_grad_apply_p2p(pol, params, adam_m, adam_v, adam_count, buf, idx + (size_t)b * B, B, B_global,
                                            adv_sums ? adv_sums + 2 * b : nullptr, &hh, peer_buffers, own_buffer, rank, world,
                                            seq, stats_out + (size_t)b * LRX_PPO_NSTATS, nonfinite_flag, error_flag,
                                            workspace, workspace_bytes, stream, &handled);
      if (rc) return rc;
      if (handled) continue;
      fused = false;
    }
    bool pushed = false;
    int rc = grad_p2p_impl(pol, params, buf, idx + (size_t)b * B, B, B_global, adv_sums ? adv_sums + 2 * b : nullptr, &hh,
                           own_buffer, seq, workspace, workspace_bytes, stream, use_push ? peer_buffers : nullptr, rank, world,
                           &pushed);
    if (rc) return rc;
    rc = apply_p2p_impl(pol, params, adam_m, adam_v, adam_count, peer_buffers, rank, world, seq, &hh,
                        stats_out + (size_t)b * LRX_PPO_NSTATS, nonfinite_flag, error_flag, scratch, scratch_bytes, stream,
                        pushed ? 1 : 0);
    if (rc) return rc;
  }
  return LRX_OK;
}

// Multi-GPU minibatch in TWO launches: the gradient kernel, then one cooperative kernel that reduces the per-CTA partials
// into this rank's exchange slot, exchanges block by block over NVLink peer memory and applies clip + Adam.
// *handled = 0 (nothing launched) for policies the fused gradient kernels do not cover: the caller then runs
// lrx_ppo_grad -> lrx_ppo_apply_p2p.
extern "C" int lrx_ppo_grad_apply_p2p(const LrxPolicyDesc* pol, float* params, float* adam_m, float* adam_v,
                                      int32_t* adam_count, const LrxBatchView* buf, const int32_t* idx, int64_t B,
                                      int64_t B_global, const double* adv_sums, const LrxPpoHyper* h,
                                      void* const* peer_buffers, void* own_buffer, int32_t rank, int32_t world,
                                      uint32_t seq, float* stats_out, int32_t* nonfinite_flag, int32_t* error_flag,
                                      void* workspace, size_t workspace_bytes, lrx_stream_t stream, int32_t* handled) {
  LRX_REQUIRE(handled, LRX_ERR_INVALID_ARG, "lrx_ppo_grad_apply_p2p: NULL handled");
  *handled = 0;
  int rc = validate_update_args(pol, buf, h);
  if (rc) return rc;
  LRX_REQUIRE(params && adam_m && adam_v && adam_count && idx && peer_buffers && own_buffer && error_flag && workspace &&
                  B > 0 && B_global >= B,
              LRX_ERR_INVALID_ARG, "lrx_ppo_grad_apply_p2p: bad argument");
  LRX_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world && seq != 0, LRX_ERR_INVALID_ARG,
              "lrx_ppo_grad_apply_p2p: bad rank / world / seq");
  LRX_REQUIRE(workspace_bytes >= lrx_ppo_workspace_bytes(pol, B), LRX_ERR_INVALID_ARG,
              "lrx_ppo_grad_apply_p2p: workspace too small");
  const int grid = (pol->n_params + 31) / 32;
  if (grid > p2p::MAX_BLOCK_FLAGS || grid > 4 * lrx_sm_count()) return LRX_OK;  // not one resident cooperative wave
  const WorkspaceTail tail = workspace_tail(pol, B);
  const int n = pol->n_params + LRX_PPO_NSTATS;
  LrxApplyArgs ap{};
  ap.params = params; ap.m = adam_m; ap.v = adam_v; ap.count = adam_count; ap.h = *h;
  ap.stats_out = stats_out; ap.nonfinite_flag = nonfinite_flag;
  ap.blk_ss = (double*)((char*)workspace + tail.scratch_off);
  ap.peers = reinterpret_cast<char* const*>(peer_buffers);
  ap.rank = rank; ap.world = world; ap.slot_b = p2p::slot_bytes(n); ap.seq = seq; ap.error_flag = error_flag;
  ap.gsum = (float*)((char*)workspace + tail.gradbuf_off);
  {
    // one flag per rank, published by the last block of the reduction (default); LRX_P2P_FUSED=1: one flag per block
    const char* e = getenv("LRX_P2P_FUSED");
    if (!(e && e[0] == '1')) {
      ap.push = (e && e[0] == '3' && world <= p2p::PUSH_MAX_RANKS) ? 3 : 2;
      ap.ticket = reinterpret_cast<unsigned*>(static_cast<char*>(own_buffer) + 64);
    }
  }
  float* slot = reinterpret_cast<float*>(static_cast<char*>(own_buffer) + p2p::HEADER + (size_t)(seq & 1u) * ap.slot_b);
  bool done = false;
  rc = lrx_ppo_grad_tensor(pol, params, buf, idx, B, B_global, adv_sums, h, slot, workspace, tail.adv_sums_off,
                           (cudaStream_t)stream, &done, &ap);
  if (rc) return rc;
  *handled = done ? 1 : 0;
  return LRX_OK;
}

extern "C" int lrx_ppo_update(const LrxPolicyDesc* pol, float* params, float* adam_m, float* adam_v,
                              int32_t* adam_count, const LrxBatchView* buf, const int32_t* idx,
                              int32_t num_batches, int64_t B, const LrxPpoHyper* h, float* stats_out,
                              int32_t* nonfinite_flag, void* workspace, size_t workspace_bytes,
                              lrx_stream_t stream) {
  return lrx_ppo_update_schedule(pol, params, adam_m, adam_v, adam_count, buf, idx, num_batches, B, h, nullptr, stats_out,
                                 nonfinite_flag, workspace, workspace_bytes, stream);
}

extern "C" int lrx_ppo_update_schedule(const LrxPolicyDesc* pol, float* params, float* adam_m, float* adam_v,
                                       int32_t* adam_count, const LrxBatchView* buf, const int32_t* idx,
                                       int32_t num_batches, int64_t B, const LrxPpoHyper* h,
                                       const float* learning_rates_host, float* stats_out, int32_t* nonfinite_flag,
                                       void* workspace, size_t workspace_bytes, lrx_stream_t stream) {
  int rc = validate_update_args(pol, buf, h);
  if (rc) return rc;
  LRX_REQUIRE(num_batches >= 0 && num_batches <= LRX_MAX_BATCHES, LRX_ERR_INVALID_ARG,
              "lrx_ppo_update: num_batches out of range [0, %d]", LRX_MAX_BATCHES);
  LRX_REQUIRE(B > 0 && idx && workspace && stats_out, LRX_ERR_INVALID_ARG, "lrx_ppo_update: bad argument");
  const size_t need = lrx_ppo_workspace_bytes(pol, B);
  LRX_REQUIRE(workspace_bytes >= need, LRX_ERR_INVALID_ARG, "lrx_ppo_update: workspace too small (%zu < %zu)",
              workspace_bytes, need);
  if (num_batches == 0) return LRX_OK;
  const WorkspaceTail tail = workspace_tail(pol, B);
  double* adv_sums = (double*)((char*)workspace + tail.adv_sums_off);
  float* gradbuf = (float*)((char*)workspace + tail.gradbuf_off);
  rc = lrx_ppo_adv_sums(buf->adv, buf->N, buf->T, idx, num_batches, B, adv_sums, stream);
  if (rc) return rc;
  LrxApplyArgs ap{};
  ap.params = params; ap.m = adam_m; ap.v = adam_v; ap.count = adam_count; ap.h = *h;
  ap.nonfinite_flag = nonfinite_flag;
  ap.blk_ss = (double*)((char*)workspace + tail.scratch_off);
  for (int b = 0; b < num_batches; ++b) {
    // default widths: one fused gradient launch + one reduce/clip/Adam launch per minibatch
    bool handled = false;
    ap.stats_out = stats_out + (size_t)b * LRX_PPO_NSTATS;
    if (learning_rates_host) ap.h.learning_rate = learning_rates_host[b];  // optax inject_hyperparams: a schedule of the count
    LRX_REQUIRE(!h->normalize_advantages || adv_sums, LRX_ERR_INVALID_ARG, "adv_sums missing");
    rc = lrx_ppo_grad_tensor(pol, params, buf, idx + (size_t)b * B, B, B, adv_sums + 2 * b, h, gradbuf, workspace,
                             tail.adv_sums_off, (cudaStream_t)stream, &handled, &ap);
    if (rc) return rc;
    if (handled) continue;
    rc = lrx_ppo_grad(pol, params, buf, idx + (size_t)b * B, B, B, adv_sums + 2 * b, h, gradbuf, workspace,
                      tail.adv_sums_off, stream);
    if (rc) return rc;
    {
      PolicyTable tab = lrx_make_table(*pol);
      rc = launch_apply(params, adam_m, adam_v, adam_count, gradbuf, pol->n_params, tab.log_std_off, ap.h,
                        stats_out + (size_t)b * LRX_PPO_NSTATS, nonfinite_flag, (cudaStream_t)stream, ap.blk_ss);
      if (rc) return rc;
      LRX_LAUNCH_CHECK();
    }
  }
  return LRX_OK;
}
