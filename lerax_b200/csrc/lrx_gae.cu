// Generalised Advantage Estimation: one HBM pass, reverse affine scan.
//
// Reference: advantage/gae.py:36-66 (via buffer/rollout.py:64-79, algorithm/ppo.py:311-315)
//     delta_t = r_t + gamma * v_{t+1} * (1-d_t) - v_t ; A_t = delta_t + gamma*lam*(1-d_t) * A_{t+1}
// The recurrence A_t = a_t + b_t * A_{t+1} is a composition of affine maps, so a strip of T
// steps is split into chunks that are scanned independently and stitched with the chunk
// summaries (a, b) -- every element is read once and written once (17 B/element).
//
//  * TN layout ([T][N], kernel-native): a warp owns a strip of 32 adjacent envs (lanes = envs,
//    128 B coalesced rows); the C warps of a block own C consecutive time chunks of L steps
//    held in registers, summaries are exchanged through shared memory.
//  * NT layout ([N][T], reference-native): a warp owns one env, lanes = 32 consecutive steps,
//    Kogge-Stone scan of the affine pairs with __shfl_down_sync.
#include "lrx_common.cuh"

template <int L>
__global__ void __launch_bounds__(512)
gae_tn_kernel(const float* __restrict__ rew, const float* __restrict__ val,
              const uint8_t* __restrict__ done, const float* __restrict__ last_val,
              float* __restrict__ adv, float* __restrict__ ret, int N, int T, float gamma,
              float lam, double* ev_sums) {
  __shared__ float sA[16][32], sP[16][32];
  __shared__ double sEv[16][4];
  const int lane = threadIdx.x, c = threadIdx.y, C = blockDim.y;
  const int n = blockIdx.x * 32 + lane;
  const bool live = n < N;
  const int nl = live ? n : N - 1;
  const float gl = gamma * lam;
  const int seg_len = C * L;
  const int num_seg = (T + seg_len - 1) / seg_len;
  const float lv = last_val[nl];
  float seg_carry = 0.f;
  double e0 = 0, e1 = 0, e2 = 0, e3 = 0;

  for (int seg = num_seg - 1; seg >= 0; --seg) {
    const int t0 = seg * seg_len + c * L;
    float r[L], v[L + 1], nd[L];
#pragma unroll
    for (int i = 0; i < L; ++i) {
      const int t = t0 + i;
      const bool ok = t < T;
      const size_t off = (size_t)(ok ? t : T - 1) * N + nl;
      r[i] = ok ? rew[off] : 0.f;
      v[i] = ok ? val[off] : lv;
      nd[i] = ok ? 1.0f - (float)done[off] : 1.0f;
    }
    v[L] = (t0 + L < T) ? val[(size_t)(t0 + L) * N + nl] : lv;

    float a = 0.f, p = 1.f;
    float A[L], P[L];
#pragma unroll
    for (int i = L - 1; i >= 0; --i) {
      const bool ok = t0 + i < T;
      const float delta = ok ? r[i] + gamma * v[i + 1] * nd[i] - v[i] : 0.f;
      const float disc = ok ? gl * nd[i] : 1.f;
      a = delta + disc * a;
      p = disc * p;
      A[i] = a;
      P[i] = p;
    }
    sA[c][lane] = a;
    sP[c][lane] = p;
    __syncthreads();
    float carry = seg_carry;
    for (int cc = C - 1; cc > c; --cc) carry = sA[cc][lane] + sP[cc][lane] * carry;
#pragma unroll
    for (int i = 0; i < L; ++i) {
      const int t = t0 + i;
      if (t < T && live) {
        const float Ai = A[i] + P[i] * carry;
        const float Ri = Ai + v[i];
        const size_t off = (size_t)t * N + n;
        adv[off] = Ai;
        ret[off] = Ri;
        e0 += Ri; e1 += (double)Ri * Ri; e2 += Ai; e3 += (double)Ai * Ai;
      }
    }
    float cr = carry;
    for (int cc = c; cc >= 0; --cc) cr = sA[cc][lane] + sP[cc][lane] * cr;
    seg_carry = cr;
    __syncthreads();
  }

  if (ev_sums) {
    e0 = warp_sum(e0); e1 = warp_sum(e1); e2 = warp_sum(e2); e3 = warp_sum(e3);
    if (lane == 0) { sEv[c][0] = e0; sEv[c][1] = e1; sEv[c][2] = e2; sEv[c][3] = e3; }
    __syncthreads();
    if (c == 0 && lane < 4) {
      double s = 0;
      for (int cc = 0; cc < C; ++cc) s += sEv[cc][lane];
      atomicAdd(ev_sums + lane, s);
    }
  }
}

// Two adjacent envs per lane (float2 / uchar2 accesses: 256-byte rows per warp, half the
// instructions); requires N even.  Same chunked affine scan as gae_tn_kernel.
template <int L>
__global__ void __launch_bounds__(512)
gae_tn2_kernel(const float* __restrict__ rew, const float* __restrict__ val,
               const uint8_t* __restrict__ done, const float* __restrict__ last_val,
               float* __restrict__ adv, float* __restrict__ ret, int N, int T, float gamma,
               float lam, double* ev_sums) {
  __shared__ float2 sA[16][32], sP[16][32];
  __shared__ double sEv[16][4];
  const int lane = threadIdx.x, c = threadIdx.y, C = blockDim.y;
  const int n = (blockIdx.x * 32 + lane) * 2;
  const bool live = n < N;
  const int nl = live ? n : N - 2;
  const float gl = gamma * lam;
  const int seg_len = C * L;
  const int num_seg = (T + seg_len - 1) / seg_len;
  const float2 lv = *reinterpret_cast<const float2*>(last_val + nl);
  float2 seg_carry = make_float2(0.f, 0.f);
  double e0 = 0, e1 = 0, e2 = 0, e3 = 0;

  for (int seg = num_seg - 1; seg >= 0; --seg) {
    const int t0 = seg * seg_len + c * L;
    float2 r[L], v[L + 1], nd[L];
#pragma unroll
    for (int i = 0; i < L; ++i) {
      const int t = t0 + i;
      const bool ok = t < T;
      const size_t off = (size_t)(ok ? t : T - 1) * N + nl;
      r[i] = ok ? *reinterpret_cast<const float2*>(rew + off) : make_float2(0.f, 0.f);
      v[i] = ok ? *reinterpret_cast<const float2*>(val + off) : lv;
      const uchar2 d = ok ? *reinterpret_cast<const uchar2*>(done + off) : make_uchar2(0, 0);
      nd[i] = make_float2(1.0f - (float)d.x, 1.0f - (float)d.y);
    }
    v[L] = (t0 + L < T) ? *reinterpret_cast<const float2*>(val + (size_t)(t0 + L) * N + nl) : lv;

    float2 a = make_float2(0.f, 0.f), p = make_float2(1.f, 1.f);
    float2 A[L], P[L];
#pragma unroll
    for (int i = L - 1; i >= 0; --i) {
      const bool ok = t0 + i < T;
      const float dx = ok ? r[i].x + gamma * v[i + 1].x * nd[i].x - v[i].x : 0.f;
      const float dy = ok ? r[i].y + gamma * v[i + 1].y * nd[i].y - v[i].y : 0.f;
      const float cx = ok ? gl * nd[i].x : 1.f, cy = ok ? gl * nd[i].y : 1.f;
      a.x = dx + cx * a.x; a.y = dy + cy * a.y;
      p.x = cx * p.x; p.y = cy * p.y;
      A[i] = a;
      P[i] = p;
    }
    sA[c][lane] = a;
    sP[c][lane] = p;
    __syncthreads();
    float2 carry = seg_carry;
    for (int cc = C - 1; cc > c; --cc) {
      const float2 qa = sA[cc][lane], qp = sP[cc][lane];
      carry.x = qa.x + qp.x * carry.x;
      carry.y = qa.y + qp.y * carry.y;
    }
    float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;  // <= 2L terms per chunk in fp32, then fp64
#pragma unroll
    for (int i = 0; i < L; ++i) {
      const int t = t0 + i;
      if (t < T && live) {
        const float2 Ai = make_float2(A[i].x + P[i].x * carry.x, A[i].y + P[i].y * carry.y);
        const float2 Ri = make_float2(Ai.x + v[i].x, Ai.y + v[i].y);
        const size_t off = (size_t)t * N + n;
        *reinterpret_cast<float2*>(adv + off) = Ai;
        *reinterpret_cast<float2*>(ret + off) = Ri;
        f0 += Ri.x + Ri.y; f1 = fmaf(Ri.x, Ri.x, fmaf(Ri.y, Ri.y, f1));
        f2 += Ai.x + Ai.y; f3 = fmaf(Ai.x, Ai.x, fmaf(Ai.y, Ai.y, f3));
      }
    }
    e0 += f0; e1 += f1; e2 += f2; e3 += f3;
    float2 cr = carry;
    for (int cc = c; cc >= 0; --cc) {
      const float2 qa = sA[cc][lane], qp = sP[cc][lane];
      cr.x = qa.x + qp.x * cr.x;
      cr.y = qa.y + qp.y * cr.y;
    }
    seg_carry = cr;
    __syncthreads();
  }

  if (ev_sums) {
    e0 = warp_sum(e0); e1 = warp_sum(e1); e2 = warp_sum(e2); e3 = warp_sum(e3);
    if (lane == 0) { sEv[c][0] = e0; sEv[c][1] = e1; sEv[c][2] = e2; sEv[c][3] = e3; }
    __syncthreads();
    if (c == 0 && lane < 4) {
      double s = 0;
      for (int cc = 0; cc < C; ++cc) s += sEv[cc][lane];
      atomicAdd(ev_sums + lane, s);
    }
  }
}

__global__ void __launch_bounds__(256)
gae_nt_kernel(const float* __restrict__ rew, const float* __restrict__ val,
              const uint8_t* __restrict__ done, const float* __restrict__ last_val,
              float* __restrict__ adv, float* __restrict__ ret, int N, int T, float gamma,
              float lam, double* ev_sums) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int env = blockIdx.x * (blockDim.x >> 5) + warp;
  if (env >= N) return;  // whole warp exits together
  const float gl = gamma * lam;
  const size_t base_off = (size_t)env * T;
  const float lv = last_val[env];
  float carry = 0.f;
  double e0 = 0, e1 = 0, e2 = 0, e3 = 0;
  for (int base = ((T - 1) / 32) * 32; base >= 0; base -= 32) {
    const int t = base + lane;
    const bool ok = t < T;
    float v = 0.f, a = 0.f, b = 1.f;
    if (ok) {
      v = val[base_off + t];
      const float vn = (t + 1 < T) ? val[base_off + t + 1] : lv;
      const float nd = 1.0f - (float)done[base_off + t];
      a = rew[base_off + t] + gamma * vn * nd - v;
      b = gl * nd;
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float a2 = __shfl_down_sync(0xffffffffu, a, o);
      const float b2 = __shfl_down_sync(0xffffffffu, b, o);
      if (lane + o < 32) {
        a = a + b * a2;
        b = b * b2;
      }
    }
    const float Ai = a + b * carry;
    if (ok) {
      const float Ri = Ai + v;
      adv[base_off + t] = Ai;
      ret[base_off + t] = Ri;
      e0 += Ri; e1 += (double)Ri * Ri; e2 += Ai; e3 += (double)Ai * Ai;
    }
    carry = __shfl_sync(0xffffffffu, Ai, 0);
  }
  if (ev_sums) {
    e0 = warp_sum(e0); e1 = warp_sum(e1); e2 = warp_sum(e2); e3 = warp_sum(e3);
    if (lane == 0) {
      atomicAdd(ev_sums + 0, e0); atomicAdd(ev_sums + 1, e1);
      atomicAdd(ev_sums + 2, e2); atomicAdd(ev_sums + 3, e3);
    }
  }
}

static int g_gae_chunk = 0;  // 0 = automatic; tools/gae_time.py sweeps it
extern "C" void lrx_debug_set_gae_chunk(int L) { g_gae_chunk = L; }

extern "C" int lrx_gae(const float* rew, const float* val, const uint8_t* done,
                       const float* last_value, float* adv, float* ret, int32_t N, int32_t T,
                       float gamma, float lam, int32_t layout, double* ev_sums,
                       lrx_stream_t stream) {
  LRX_REQUIRE(N >= 0 && T >= 0, LRX_ERR_INVALID_ARG, "lrx_gae: negative N or T");
  LRX_REQUIRE(layout == LRX_LAYOUT_TN || layout == LRX_LAYOUT_NT, LRX_ERR_INVALID_ARG,
              "lrx_gae: unknown layout %d", layout);
  cudaStream_t s = (cudaStream_t)stream;
  if (ev_sums) LRX_CUDA_CHECK(cudaMemsetAsync(ev_sums, 0, 4 * sizeof(double), s));
  if (N == 0 || T == 0) return LRX_OK;
  LRX_REQUIRE(rew && val && done && last_value && adv && ret, LRX_ERR_INVALID_ARG, "lrx_gae: NULL pointer");
  if (layout == LRX_LAYOUT_TN) {
    const int blocks = (N + 31) / 32;
    const int Lsel = g_gae_chunk > 0 ? g_gae_chunk : 24;
    if (T >= 64 && (N % 2) == 0 && N >= 64 && (Lsel == 8 || Lsel == 28 || Lsel == 24)) {
      const int blocks2 = (N / 2 + 31) / 32;
      if (Lsel == 24) {
        int C = (T + 3) / 4;
        if (C > 16) C = 16;
        gae_tn2_kernel<4><<<blocks2, dim3(32, C), 0, s>>>(rew, val, done, last_value, adv, ret, N, T, gamma, lam, ev_sums);
      } else {
        int C = (T + 7) / 8;
        if (C > 16) C = 16;
        gae_tn2_kernel<8><<<blocks2, dim3(32, C), 0, s>>>(rew, val, done, last_value, adv, ret, N, T, gamma, lam, ev_sums);
      }
    } else if (T >= 64 && Lsel == 16) {
      int C = (T + 15) / 16;
      if (C > 16) C = 16;
      gae_tn_kernel<16><<<blocks, dim3(32, C), 0, s>>>(rew, val, done, last_value, adv, ret, N, T, gamma, lam, ev_sums);
    } else if (T >= 64 && (Lsel == 8 || Lsel == 18)) {
      int C = (T + 7) / 8;
      if (C > 16) C = 16;
      gae_tn_kernel<8><<<blocks, dim3(32, C), 0, s>>>(rew, val, done, last_value, adv, ret, N, T, gamma, lam, ev_sums);
    } else {
      int C = (T + 3) / 4;
      if (C > 16) C = 16;
      gae_tn_kernel<4><<<blocks, dim3(32, C), 0, s>>>(rew, val, done, last_value, adv, ret, N, T, gamma, lam, ev_sums);
    }
  } else {
    const int warps = 8;
    gae_nt_kernel<<<(N + warps - 1) / warps, warps * 32, 0, s>>>(rew, val, done, last_value, adv, ret, N, T, gamma, lam, ev_sums);
  }
  LRX_LAUNCH_CHECK();
  return LRX_OK;
}

// ================================================================================ other estimators
// BootstrappedReturn (advantage/bootstrapped.py:12-79) and NStepReturn (advantage/n_step.py:30-70): the two
// other AbstractAdvantageEstimators of the reference.  They are not on the PPO path
// (buffer/rollout.py:70 always builds a GAE); one thread per env (bootstrapped: a serial reverse scan) or per
// element (n-step: n Horner steps), unfused multiplies and adds so the result equals the NumPy oracle bit for bit.
__device__ __forceinline__ size_t est_index(int layout, int N, int T, int n, int t) {
  return layout == LRX_LAYOUT_TN ? (size_t)t * N + n : (size_t)n * T + t;
}

__global__ void bootstrapped_return_kernel(const float* __restrict__ rew, const float* __restrict__ val,
                                           const uint8_t* __restrict__ done, const float* __restrict__ last_value,
                                           float* __restrict__ adv, float* __restrict__ ret, int N, int T, float gamma,
                                           int layout) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float carry = last_value[n];
  for (int t = T - 1; t >= 0; --t) {
    const size_t i = est_index(layout, N, T, n, t);
    const float not_done = done[i] ? 0.0f : 1.0f;
    carry = __fadd_rn(rew[i], __fmul_rn(__fmul_rn(gamma, not_done), carry));  // reward + gamma * discount * carry
    ret[i] = carry;
    adv[i] = __fsub_rn(carry, val[i]);
  }
}

__global__ void n_step_return_kernel(const float* __restrict__ rew, const float* __restrict__ val,
                                     const uint8_t* __restrict__ done, const float* __restrict__ last_value,
                                     float* __restrict__ adv, float* __restrict__ ret, int N, int T, float gamma, int nstep,
                                     int layout) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)N * T) return;
  // consecutive threads walk the contiguous axis of the layout
  const int n = layout == LRX_LAYOUT_TN ? (int)(e % N) : (int)(e / T);
  const int t = layout == LRX_LAYOUT_TN ? (int)(e / N) : (int)(e % T);
  // bootstrap value n steps ahead (last_value past the segment), then Horner from k = n-1 down to 0 over the
  // zero-reward / unit-discount padding of n_step.py:43-52
  float carry = (t + nstep < T) ? val[est_index(layout, N, T, n, t + nstep)] : last_value[n];
  for (int k = nstep - 1; k >= 0; --k) {
    const int tk = t + k;
    if (tk < T) {
      const size_t i = est_index(layout, N, T, n, tk);
      const float disc = __fmul_rn(gamma, done[i] ? 0.0f : 1.0f);
      carry = __fadd_rn(rew[i], __fmul_rn(disc, carry));
    }  // padding: 0 + 1 * carry
  }
  const size_t i0 = est_index(layout, N, T, n, t);
  ret[i0] = carry;
  adv[i0] = __fsub_rn(carry, val[i0]);
}

extern "C" int lrx_returns(const float* rew, const float* val, const uint8_t* done, const float* last_value, float* adv,
                           float* ret, int32_t N, int32_t T, float gamma, int32_t n_step, int32_t layout,
                           lrx_stream_t stream) {
  LRX_REQUIRE(N >= 0 && T >= 0, LRX_ERR_INVALID_ARG, "lrx_returns: negative size");
  LRX_REQUIRE(n_step >= 0, LRX_ERR_INVALID_ARG, "lrx_returns: n must be at least 1 (0 selects BootstrappedReturn)");
  LRX_REQUIRE(layout == LRX_LAYOUT_TN || layout == LRX_LAYOUT_NT, LRX_ERR_INVALID_ARG, "lrx_returns: unknown layout %d",
              layout);
  if (N == 0 || T == 0) return LRX_OK;
  LRX_REQUIRE(rew && val && done && last_value && adv && ret, LRX_ERR_INVALID_ARG, "lrx_returns: NULL pointer");
  cudaStream_t s = (cudaStream_t)stream;
  if (n_step == 0) {
    bootstrapped_return_kernel<<<(N + 127) / 128, 128, 0, s>>>(rew, val, done, last_value, adv, ret, N, T, gamma, layout);
  } else {
    const long long total = (long long)N * T;
    n_step_return_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(rew, val, done, last_value, adv, ret, N, T, gamma,
                                                                         n_step, layout);
  }
  LRX_LAUNCH_CHECK();
  return LRX_OK;
}
