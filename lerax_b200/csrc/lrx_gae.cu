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

// Explained-variance sums (algorithm/ppo.py:523-528) without atomics: every block stores its four partial sums in
// ev[4 + 4 * block], the last block to arrive (ticket at ev[LRX_GAE_EV_DOUBLES - 1]) adds them in block order and
// re-arms the ticket, so the result is bitwise reproducible and no memset launch precedes the kernel.
__device__ __forceinline__ void ev_block_finish(double* ev, const double (*sEv)[4], int C, int tid) {
  __shared__ bool last;
  __shared__ double red[64][4];
  if (tid < 4) {
    double s = 0;
    for (int cc = 0; cc < C; ++cc) s += sEv[cc][tid];
    ev[4 + 4 * (size_t)blockIdx.x + tid] = s;
    __threadfence();  // only the partial sums need to be visible before the ticket; the adv / ret stores drain on their own
  }
  __syncthreads();
  unsigned long long* ticket = reinterpret_cast<unsigned long long*>(ev + LRX_GAE_EV_DOUBLES - 1);
  if (tid == 0) last = atomicAdd(ticket, 1ull) == (unsigned long long)gridDim.x - 1;
  __syncthreads();
  if (!last) return;  // block-uniform
  __threadfence();
  // fixed order: `parts` threads per quantity each add every parts-th block partial, then the parts in order
  const int nthreads = blockDim.x * blockDim.y;
  const int parts = nthreads >= 256 ? 64 : nthreads / 4;
  if (tid < 4 * parts) {
    const int k = tid & 3, part = tid >> 2;
    double s = 0;
#pragma unroll 8
    for (unsigned b = part; b < gridDim.x; b += parts) s += __ldcg(ev + 4 + 4 * (size_t)b + k);
    red[part][k] = s;
  }
  __syncthreads();
  if (tid < 4) {
    double s = 0;
    for (int p = 0; p < parts; ++p) s += red[p][tid];
    ev[tid] = s;
  }
  if (tid == 0) *ticket = 0ull;
}

// EPL adjacent envs per lane (EPL = 1: 128-byte rows per warp; 2: float2 / uchar2, 256-byte rows; 4: float4 /
// uchar4, 512-byte rows), C warps own C consecutive chunks of L steps held in registers; chunk summaries (a, p) of
// the affine maps are exchanged through shared memory.  Strips of 32 * EPL envs are walked with a grid stride so the
// launch can be sized to one resident wave.  EPL > 1 needs N % EPL == 0 and EPL*4-byte aligned planes.
template <int EPL> struct GaeVec;
template <> struct GaeVec<1> { using F = float; using U = uint8_t; };
template <> struct GaeVec<2> { using F = float2; using U = uchar2; };
template <> struct GaeVec<4> { using F = float4; using U = uchar4; };

template <int L, int EPL, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
gae_tn_kernel(const float* __restrict__ rew, const float* __restrict__ val, const uint8_t* __restrict__ done,
              const float* __restrict__ last_val, float* __restrict__ adv, float* __restrict__ ret, int N, int T,
              float gamma, float lam, double* ev_sums) {
  using F = typename GaeVec<EPL>::F;
  using U = typename GaeVec<EPL>::U;
  __shared__ float sA[16][32 * EPL], sP[16][32 * EPL];
  __shared__ double sEv[16][4];
  const int lane = threadIdx.x, c = threadIdx.y, C = blockDim.y;
  const float gl = gamma * lam;
  const int seg_len = C * L;
  const int num_seg = (T + seg_len - 1) / seg_len;
  const int n_strips = (N + 32 * EPL - 1) / (32 * EPL);
  double e0 = 0, e1 = 0, e2 = 0, e3 = 0;

  for (int strip = blockIdx.x; strip < n_strips; strip += gridDim.x) {
    const int n = (strip * 32 + lane) * EPL;
    const bool live = n < N;
    const int nl = live ? n : N - EPL;
    float lv[EPL];
    {
      const F x = *reinterpret_cast<const F*>(last_val + nl);
#pragma unroll
      for (int e = 0; e < EPL; ++e) lv[e] = reinterpret_cast<const float*>(&x)[e];
    }
    float seg_carry[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) seg_carry[e] = 0.f;

    for (int seg = num_seg - 1; seg >= 0; --seg) {
      const int t0 = seg * seg_len + c * L;
      float r[L][EPL], v[L + 1][EPL], nd[L][EPL];
#pragma unroll
      for (int i = 0; i < L; ++i) {
        const int t = t0 + i;
        const bool ok = t < T;
        const size_t off = (size_t)(ok ? t : T - 1) * N + nl;
        const F rr = *reinterpret_cast<const F*>(rew + off);
        const F vv = *reinterpret_cast<const F*>(val + off);
        const U dd = *reinterpret_cast<const U*>(done + off);
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          r[i][e] = ok ? reinterpret_cast<const float*>(&rr)[e] : 0.f;
          v[i][e] = ok ? reinterpret_cast<const float*>(&vv)[e] : lv[e];
          nd[i][e] = ok ? 1.0f - (float)reinterpret_cast<const uint8_t*>(&dd)[e] : 1.0f;
        }
      }
      {
        const bool ok = t0 + L < T;
        const F vv = *reinterpret_cast<const F*>(val + (size_t)(ok ? t0 + L : T - 1) * N + nl);
#pragma unroll
        for (int e = 0; e < EPL; ++e) v[L][e] = ok ? reinterpret_cast<const float*>(&vv)[e] : lv[e];
      }
      float a[EPL], p[EPL], A[L][EPL], P[L][EPL];
#pragma unroll
      for (int e = 0; e < EPL; ++e) { a[e] = 0.f; p[e] = 1.f; }
#pragma unroll
      for (int i = L - 1; i >= 0; --i) {
        const bool ok = t0 + i < T;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          const float delta = ok ? r[i][e] + gamma * v[i + 1][e] * nd[i][e] - v[i][e] : 0.f;
          const float disc = ok ? gl * nd[i][e] : 1.f;
          a[e] = delta + disc * a[e];
          p[e] = disc * p[e];
          A[i][e] = a[e];
          P[i][e] = p[e];
        }
      }
#pragma unroll
      for (int e = 0; e < EPL; ++e) { sA[c][lane * EPL + e] = a[e]; sP[c][lane * EPL + e] = p[e]; }
      __syncthreads();
      float carry[EPL];
#pragma unroll
      for (int e = 0; e < EPL; ++e) carry[e] = seg_carry[e];
      for (int cc = C - 1; cc > c; --cc) {
#pragma unroll
        for (int e = 0; e < EPL; ++e) carry[e] = sA[cc][lane * EPL + e] + sP[cc][lane * EPL + e] * carry[e];
      }
      float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;  // <= EPL * L terms per chunk in fp32, then fp64
#pragma unroll
      for (int i = 0; i < L; ++i) {
        const int t = t0 + i;
        if (t < T && live) {
          F Ao, Ro;
#pragma unroll
          for (int e = 0; e < EPL; ++e) {
            const float Ai = A[i][e] + P[i][e] * carry[e];
            const float Ri = Ai + v[i][e];
            reinterpret_cast<float*>(&Ao)[e] = Ai;
            reinterpret_cast<float*>(&Ro)[e] = Ri;
            f0 += Ri; f1 = fmaf(Ri, Ri, f1); f2 += Ai; f3 = fmaf(Ai, Ai, f3);
          }
          const size_t off = (size_t)t * N + n;
          *reinterpret_cast<F*>(adv + off) = Ao;
          *reinterpret_cast<F*>(ret + off) = Ro;
        }
      }
      e0 += f0; e1 += f1; e2 += f2; e3 += f3;
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        float cr = carry[e];
        for (int cc = c; cc >= 0; --cc) cr = sA[cc][lane * EPL + e] + sP[cc][lane * EPL + e] * cr;
        seg_carry[e] = cr;
      }
      __syncthreads();
    }
  }

  if (ev_sums) {
    e0 = warp_sum(e0); e1 = warp_sum(e1); e2 = warp_sum(e2); e3 = warp_sum(e3);
    if (lane == 0) { sEv[c][0] = e0; sEv[c][1] = e1; sEv[c][2] = e2; sEv[c][3] = e3; }
    __syncthreads();
    ev_block_finish(ev_sums, sEv, C, threadIdx.y * 32 + threadIdx.x);
  }
}

// Streaming variant for wide buffers (N >= a few thousand env strips): ONE WARP PER ENV STRIP walks all T steps of its
// 32 * EPL envs from t = T-1 down to 0 with the recurrence carried in registers -- the reference's own order of
// operations (gae.py:47-66), no chunk summaries, no shared memory, no block barrier.  The loads are software pipelined
// in groups of G steps: group k+1's rew / val / done rows are requested before group k is scanned, so a warp keeps
// 2 * G * EPL * 9 bytes per lane in flight and ~7 warps per SM cover the HBM latency.  ncu on the chunked kernel at
// config 2 (profiles/r2_gae.md): 66 thread instructions per element, issue slots 60 % busy at 39 % occupancy -- that
// kernel is instruction bound below ~10^7 elements; this one executes ~14 per element.
template <int EPL, int G, bool CS>
__global__ void __launch_bounds__(256)
gae_stream_kernel(const float* __restrict__ rew, const float* __restrict__ val, const uint8_t* __restrict__ done,
                  const float* __restrict__ last_val, float* __restrict__ adv, float* __restrict__ ret, int N, int T,
                  float gamma, float lam, double* ev_sums) {
  using F = typename GaeVec<EPL>::F;
  using U = typename GaeVec<EPL>::U;
  __shared__ double sEv[16][4];
  const int lane = threadIdx.x, w = threadIdx.y, W = blockDim.y;
  const float gl = gamma * lam;
  const int n_strips = (N + 32 * EPL - 1) / (32 * EPL);
  const int ng = (T + G - 1) / G;
  double e0 = 0, e1 = 0, e2 = 0, e3 = 0;
  struct Group { F r[G], v[G]; U d[G]; };

  for (int strip = blockIdx.x * W + w; strip < n_strips; strip += gridDim.x * W) {  // warp-uniform
    const int n = (strip * 32 + lane) * EPL;
    const bool live = n < N;
    const int nl = live ? n : N - EPL;
    auto load = [&](Group& gq, int k) {  // group k covers t in [T - (k+1) G, T - k G); rows below t = 0 re-read row 0
      const int base = T - (k + 1) * G;
#pragma unroll
      for (int i = 0; i < G; ++i) {
        const int t = base + i;
        const size_t off = (size_t)(t > 0 ? t : 0) * N + nl;
        if (CS) {
          gq.r[i] = __ldcs(reinterpret_cast<const F*>(rew + off));
          gq.v[i] = __ldcs(reinterpret_cast<const F*>(val + off));
          gq.d[i] = __ldcs(reinterpret_cast<const U*>(done + off));
        } else {
          gq.r[i] = *reinterpret_cast<const F*>(rew + off);
          gq.v[i] = *reinterpret_cast<const F*>(val + off);
          gq.d[i] = *reinterpret_cast<const U*>(done + off);
        }
      }
    };
    float vn[EPL], A[EPL];  // value of step t + 1 (the bootstrap value past the end) and advantage of step t + 1
    {
      const F x = *reinterpret_cast<const F*>(last_val + nl);
#pragma unroll
      for (int e = 0; e < EPL; ++e) { vn[e] = reinterpret_cast<const float*>(&x)[e]; A[e] = 0.f; }
    }
    auto scan = [&](const Group& gq, int k) {
      const int base = T - (k + 1) * G;
      float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;  // <= G * EPL terms in fp32, then fp64
#pragma unroll
      for (int i = G - 1; i >= 0; --i) {
        const int t = base + i;
        if (t >= 0) {  // warp-uniform
          F Ao, Ro;
#pragma unroll
          for (int e = 0; e < EPL; ++e) {
            const float r = reinterpret_cast<const float*>(&gq.r[i])[e], v = reinterpret_cast<const float*>(&gq.v[i])[e];
            const float nd = 1.0f - (float)reinterpret_cast<const uint8_t*>(&gq.d[i])[e];
            const float delta = r + gamma * vn[e] * nd - v;
            A[e] = delta + gl * nd * A[e];
            vn[e] = v;
            const float Ri = A[e] + v;
            reinterpret_cast<float*>(&Ao)[e] = A[e];
            reinterpret_cast<float*>(&Ro)[e] = Ri;
            f0 += Ri; f1 = fmaf(Ri, Ri, f1); f2 += A[e]; f3 = fmaf(A[e], A[e], f3);
          }
          if (live) {
            const size_t off = (size_t)t * N + n;
            if (CS) {
              __stcs(reinterpret_cast<F*>(adv + off), Ao);
              __stcs(reinterpret_cast<F*>(ret + off), Ro);
            } else {
              *reinterpret_cast<F*>(adv + off) = Ao;
              *reinterpret_cast<F*>(ret + off) = Ro;
            }
          }
        }
      }
      if (live) { e0 += f0; e1 += f1; e2 += f2; e3 += f3; }
    };
    Group ga, gb;
    load(ga, 0);
    for (int k = 0; k < ng; k += 2) {
      if (k + 1 < ng) load(gb, k + 1);
      scan(ga, k);
      if (k + 1 < ng) {
        if (k + 2 < ng) load(ga, k + 2);
        scan(gb, k + 1);
      }
    }
  }

  if (ev_sums) {
    e0 = warp_sum(e0); e1 = warp_sum(e1); e2 = warp_sum(e2); e3 = warp_sum(e3);
    if (lane == 0) { sEv[w][0] = e0; sEv[w][1] = e1; sEv[w][2] = e2; sEv[w][3] = e3; }
    __syncthreads();
    ev_block_finish(ev_sums, sEv, W, threadIdx.y * 32 + threadIdx.x);
  }
}

__global__ void __launch_bounds__(256)
gae_nt_kernel(const float* __restrict__ rew, const float* __restrict__ val,
              const uint8_t* __restrict__ done, const float* __restrict__ last_val,
              float* __restrict__ adv, float* __restrict__ ret, int N, int T, float gamma,
              float lam, double* ev_sums) {
  __shared__ double sEv[8][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const float gl = gamma * lam;
  double e0 = 0, e1 = 0, e2 = 0, e3 = 0;
  for (int env = blockIdx.x * warps + warp; env < N; env += gridDim.x * warps) {  // warp-uniform
    const size_t base_off = (size_t)env * T;
    const float lv = last_val[env];
    float carry = 0.f;
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
  }
  if (ev_sums) {
    e0 = warp_sum(e0); e1 = warp_sum(e1); e2 = warp_sum(e2); e3 = warp_sum(e3);
    if (lane == 0) { sEv[warp][0] = e0; sEv[warp][1] = e1; sEv[warp][2] = e2; sEv[warp][3] = e3; }
    __syncthreads();
    ev_block_finish(ev_sums, sEv, warps, threadIdx.x);
  }
}

// tools/gae_time.py sweeps the launch shape: knob = MINB * 1000000 + EPL * 10000 + L * 100 + C (0 = automatic):
// MINB resident blocks per SM asked of the compiler (register cap), EPL envs per lane, L steps per thread, C warps.
static int g_gae_knob = 0;
extern "C" void lrx_debug_set_gae_chunk(int knob) { g_gae_knob = knob; }

template <int L, int EPL, int MAXT, int MINB>
static void launch_gae_tn(int C, int grid, cudaStream_t s, const float* rew, const float* val, const uint8_t* done,
                          const float* last_value, float* adv, float* ret, int N, int T, float gamma, float lam,
                          double* ev_sums) {
  gae_tn_kernel<L, EPL, MAXT, MINB><<<grid, dim3(32, C), 0, s>>>(rew, val, done, last_value, adv, ret, N, T, gamma, lam, ev_sums);
}

extern "C" int lrx_gae(const float* rew, const float* val, const uint8_t* done,
                       const float* last_value, float* adv, float* ret, int32_t N, int32_t T,
                       float gamma, float lam, int32_t layout, double* ev_sums,
                       lrx_stream_t stream) {
  LRX_REQUIRE(N >= 0 && T >= 0, LRX_ERR_INVALID_ARG, "lrx_gae: negative N or T");
  LRX_REQUIRE(layout == LRX_LAYOUT_TN || layout == LRX_LAYOUT_NT, LRX_ERR_INVALID_ARG,
              "lrx_gae: unknown layout %d", layout);
  cudaStream_t s = (cudaStream_t)stream;
  if (N == 0 || T == 0) {
    if (ev_sums) LRX_CUDA_CHECK(cudaMemsetAsync(ev_sums, 0, 4 * sizeof(double), s));
    return LRX_OK;
  }
  LRX_REQUIRE(rew && val && done && last_value && adv && ret, LRX_ERR_INVALID_ARG, "lrx_gae: NULL pointer");
  if (layout == LRX_LAYOUT_TN) {
    // vector width: all planes (and every row t * N of them) must be aligned to the access size
    auto aligned = [&](int epl) {
      const uintptr_t m = (uintptr_t)rew | (uintptr_t)val | (uintptr_t)adv | (uintptr_t)ret | (uintptr_t)last_value;
      return N % epl == 0 && (m & (uintptr_t)(4 * epl - 1)) == 0 && ((uintptr_t)done & (uintptr_t)(epl - 1)) == 0;
    };
    // tools/gae_time.py sweep on B200 (profiles/r2_gae.md): two envs per lane, four steps per thread, <= 48 registers (five
    // 256-thread blocks per SM); four warps per block, eight once the buffer is far larger than one wave
    int epl = 2, L = 4, C = ((long long)N * T >= (1ll << 26)) ? 8 : 4, minb = 5;
    if (g_gae_knob > 0) {
      minb = g_gae_knob / 1000000; epl = (g_gae_knob / 10000) % 100; L = (g_gae_knob / 100) % 100; C = g_gae_knob % 100;
    }
    if (N < 64 * epl) epl = 1;
    while (epl > 1 && !aligned(epl)) epl >>= 1;
    // Streaming kernel (one warp per strip, the whole horizon through a register pipeline): chosen when the strips alone
    // fill the GPU with several warps per SM; knob < 0 forces it: -(CS * 1000000 + W * 10000 + EPL * 100 + G), CS = 1:
    // plain instead of evict-first loads / stores.
    {
      // tools/gae_time.py on B200 (profiles/r2_gae.md): plain loads / stores beat the evict-first hints by 4 - 10 %; two envs
      // per lane and 16-step groups up to ~3 x 10^5 envs, four envs per lane and 8-step groups beyond
      int sw = 4, se = 2, sg = 16, plain = 1;
      bool want = g_gae_knob == 0 && epl >= 2 && (N + 63) / 64 >= 4 * lrx_sm_count();
      if (want && aligned(4) && (N + 127) / 128 >= 16 * lrx_sm_count()) { se = 4; sg = 8; }
      if (g_gae_knob < 0) {
        const int k = -g_gae_knob;
        plain = k / 1000000; sw = (k / 10000) % 100; se = (k / 100) % 100; sg = k % 100;
        want = se >= 1 && se <= 4 && aligned(se) && N >= 32 * se;
        if (!want) { minb = 0; epl = 2; L = 4; C = 4; if (N < 64 * epl) epl = 1; while (epl > 1 && !aligned(epl)) epl >>= 1; }
      }
      if (want) {
        const int n_strips = (N + 32 * se - 1) / (32 * se);
        const int blocks = (n_strips + sw - 1) / sw;
        int per_sm = 1;
        bool ok = true;
#define LRX_GAE_STREAM(EE, GG, CC)                                                                                     \
  if (se == EE && sg == GG && plain == (CC ? 0 : 1)) {                                                                \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gae_stream_kernel<EE, GG, CC>, 32 * sw, 0);                \
    int grid = blocks < per_sm * lrx_sm_count() ? blocks : per_sm * lrx_sm_count();                                   \
    if (grid > LRX_GAE_EV_BLOCKS) grid = LRX_GAE_EV_BLOCKS;                                                           \
    gae_stream_kernel<EE, GG, CC><<<grid, dim3(32, sw), 0, s>>>(rew, val, done, last_value, adv, ret, N, T, gamma,    \
                                                                lam, ev_sums);                                         \
  } else
        LRX_GAE_STREAM(2, 16, true) LRX_GAE_STREAM(2, 8, true) LRX_GAE_STREAM(1, 16, true) LRX_GAE_STREAM(4, 8, true)
        LRX_GAE_STREAM(2, 16, false) LRX_GAE_STREAM(4, 8, false) LRX_GAE_STREAM(2, 12, true) { ok = false; }
#undef LRX_GAE_STREAM
        LRX_REQUIRE(ok && sw >= 1 && sw <= 8, LRX_ERR_INVALID_ARG, "lrx_gae: unsupported streaming shape W=%d EPL=%d G=%d", sw, se, sg);
        LRX_LAUNCH_CHECK();
        return LRX_OK;
      }
    }
    if (C > 16) C = 16;
    while (C > 1 && (C - 1) * L >= T) --C;  // no warp without a step
    const int n_strips = (N + 32 * epl - 1) / (32 * epl);
    const int grid = n_strips < LRX_GAE_EV_BLOCKS ? n_strips : LRX_GAE_EV_BLOCKS;
    bool launched = false;
#define LRX_GAE_CASE(LL, EE, MT, MB)                                                                                  \
  if (!launched && L == LL && epl == EE && C * 32 <= MT && (minb == MB || any_minb)) {                                 \
    launch_gae_tn<LL, EE, MT, MB>(C, grid, s, rew, val, done, last_value, adv, ret, N, T, gamma, lam, ev_sums);        \
    launched = true;                                                                                                  \
  }
    for (int pass = 0; pass < 2 && !launched; ++pass) {  // the requested register cap first, then any instantiation
      const bool any_minb = pass == 1 && (g_gae_knob == 0 || minb == 0);
      LRX_GAE_CASE(4, 2, 256, 5) LRX_GAE_CASE(4, 2, 256, 4) LRX_GAE_CASE(4, 2, 256, 3) LRX_GAE_CASE(4, 2, 512, 2)
      LRX_GAE_CASE(2, 2, 256, 6) LRX_GAE_CASE(2, 2, 512, 3) LRX_GAE_CASE(8, 2, 256, 3) LRX_GAE_CASE(8, 2, 512, 1)
      LRX_GAE_CASE(4, 1, 256, 6) LRX_GAE_CASE(4, 1, 512, 2) LRX_GAE_CASE(2, 1, 256, 8) LRX_GAE_CASE(8, 1, 256, 4)
      LRX_GAE_CASE(4, 4, 256, 3) LRX_GAE_CASE(4, 4, 256, 2) LRX_GAE_CASE(2, 4, 256, 4) LRX_GAE_CASE(2, 4, 512, 2)
    }
    LRX_REQUIRE(launched, LRX_ERR_INVALID_ARG, "lrx_gae: unsupported launch shape L=%d epl=%d C=%d minb=%d", L, epl, C, minb);
#undef LRX_GAE_CASE
  } else {
    const int warps = 8;
    int grid = (N + warps - 1) / warps;
    if (grid > LRX_GAE_EV_BLOCKS) grid = LRX_GAE_EV_BLOCKS;
    gae_nt_kernel<<<grid, warps * 32, 0, s>>>(rew, val, done, last_value, adv, ret, N, T, gamma, lam, ev_sums);
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
