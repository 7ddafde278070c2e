// Self-test of the tcgen05 building blocks (lrx_umma.cuh): one CTA computes D = A * B^T with
// kind::tf32 MMAs (optionally the 3xTF32 split) from operands staged in the canonical shared
// memory layout, in K-major or MN-major form, and dumps the raw TMEM accumulator
// [128 lanes][N columns] so the test can also check the lane mapping of M = 64.
#include <stdlib.h>

#include "lrx_common.cuh"
#include "lrx_umma.cuh"

struct SelftestOperand {
  int mn;        // 0: K-major core matrices (8 mn x 16 B of k), 1: MN-major (8 k x 16 B of mn)
  int s_mn;      // storage byte stride between core-matrix groups along MN
  int s_k;       // storage byte stride between core-matrix groups along K
  int lbo, sbo;  // descriptor fields (bytes)
  int kstep;     // start-address advance per MMA (K = 8), bytes
};
struct SelftestCfg {
  SelftestOperand a, b;
};

__device__ __forceinline__ uint32_t selftest_off(const SelftestOperand& o, int mn, int k) {
  return o.mn ? (uint32_t)((mn & 3) * 4 + (mn >> 2) * o.s_mn + (k & 7) * 16 + (k >> 3) * o.s_k)
              : (uint32_t)((mn & 7) * 16 + (mn >> 3) * o.s_mn + (k & 3) * 4 + (k >> 2) * o.s_k);
}

__global__ void __launch_bounds__(128)
umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ dump,
                     int M, int N, int K, SelftestCfg cfg, int split, uint32_t bytesA, uint32_t bytesB) {  // NOLINT
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;

  uint8_t* a_hi = smem;
  uint8_t* a_lo = a_hi + bytesA;
  uint8_t* b_hi = a_lo + bytesA;
  uint8_t* b_lo = b_hi + bytesB;
  const bool raw_a = cfg.a.mn & 2, raw_b = cfg.b.mn & 2;  // probe mode: fill with the word index
  const uint32_t lt_a = (cfg.a.mn >> 4) & 7, lt_b = (cfg.b.mn >> 4) & 7;  // descriptor layout_type (probe)
  cfg.a.mn &= 1;
  cfg.b.mn &= 1;
  if (raw_a) {
    for (uint32_t w = tid; w < bytesA / 4; w += blockDim.x) {
      ((float*)a_hi)[w] = umma::tf32_hi((float)w);
      ((float*)a_lo)[w] = umma::tf32_lo((float)w);
    }
  } else {
    for (int i = tid; i < M * K; i += blockDim.x) {
      const int m = i / K, k = i % K;
      const float x = A[i];
      const uint32_t off = selftest_off(cfg.a, m, k);
      *(float*)(a_hi + off) = umma::tf32_hi(x);
      *(float*)(a_lo + off) = umma::tf32_lo(x);
    }
  }
  if (raw_b) {
    for (uint32_t w = tid; w < bytesB / 4; w += blockDim.x) {
      ((float*)b_hi)[w] = umma::tf32_hi((float)w);
      ((float*)b_lo)[w] = umma::tf32_lo((float)w);
    }
  } else {
    for (int i = tid; i < N * K; i += blockDim.x) {
      const int n = i / K, k = i % K;
      const float x = B[i];
      const uint32_t off = selftest_off(cfg.b, n, k);
      *(float*)(b_hi + off) = umma::tf32_hi(x);
      *(float*)(b_lo + off) = umma::tf32_lo(x);
    }
  }
  if (tid == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 256);
  umma::fence_async_smem();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tbase = tmem_slot;

  // clear the accumulator region through tcgen05.st so untouched lanes read back as zero
  {
    float z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0.f;
    for (int c = 0; c < 256; c += 16) umma::tmem_st16(umma::tmem_addr(tbase, warp * 32, c), z);
    umma::tmem_st_wait();
  }
  umma::tc_fence_before();
  __syncthreads();

  if (tid == 0) {
    umma::tc_fence_after();
    const uint32_t idesc = umma::idesc_tf32(M, N, cfg.a.mn, cfg.b.mn);
    const uint32_t ah = umma::smem_u32(a_hi), al = umma::smem_u32(a_lo);
    const uint32_t bh = umma::smem_u32(b_hi), bl = umma::smem_u32(b_lo);
    uint32_t acc = 0;
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint32_t ao = ks * cfg.a.kstep, bo = ks * cfg.b.kstep;
      if (split) {
        umma::mma_tf32(tbase, umma::smem_desc(al + ao, cfg.a.lbo, cfg.a.sbo, lt_a), umma::smem_desc(bh + bo, cfg.b.lbo, cfg.b.sbo, lt_b), idesc, acc);
        umma::mma_tf32(tbase, umma::smem_desc(ah + ao, cfg.a.lbo, cfg.a.sbo, lt_a), umma::smem_desc(bl + bo, cfg.b.lbo, cfg.b.sbo, lt_b), idesc, 1u);
        acc = 1u;
      }
      umma::mma_tf32(tbase, umma::smem_desc(ah + ao, cfg.a.lbo, cfg.a.sbo, lt_a), umma::smem_desc(bh + bo, cfg.b.lbo, cfg.b.sbo, lt_b), idesc, acc);
      acc = 1u;
    }
    umma::mma_commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::tc_fence_after();

  for (int c = 0; c < N; c += 8) {
    float v[8];
    umma::tmem_ld8(umma::tmem_addr(tbase, warp * 32, c), v);
    umma::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) dump[(size_t)tid * N + c + i] = v[i];
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 256);
}

static uint32_t selftest_extent(SelftestOperand o, int MN, int K) {
  // one past the last byte any element touches
  o.mn &= 1;
  uint32_t mx = 0;
  for (int mn = 0; mn < MN; ++mn)
    for (int k = 0; k < K; ++k) {
      uint32_t off = o.mn ? (uint32_t)((mn & 3) * 4 + (mn >> 2) * o.s_mn + (k & 7) * 16 + (k >> 3) * o.s_k)
                          : (uint32_t)((mn & 7) * 16 + (mn >> 3) * o.s_mn + (k & 3) * 4 + (k >> 2) * o.s_k);
      if (off + 4 > mx) mx = off + 4;
    }
  return (mx + 1023u) / 1024u * 1024u;
}

// D = A[M][K] * B[N][K]^T on one SM's tensor core; dump receives the raw accumulator
// [128][N] (TMEM lane major).  cfg_host: 12 ints = {mn, s_mn, s_k, lbo, sbo, kstep} for A then B
// (storage strides and descriptor fields in bytes).  Test-only entry point (tests/test_gpu_umma.py).
extern "C" int lrx_debug_umma_gemm(const float* A, const float* B, float* dump, int32_t M, int32_t N,
                                   int32_t K, const int32_t* cfg_host, int32_t split, lrx_stream_t stream) {
  LRX_REQUIRE(A && B && dump && cfg_host, LRX_ERR_INVALID_ARG, "lrx_debug_umma_gemm: NULL pointer");
  LRX_REQUIRE((M == 64 || M == 128) && N >= 8 && N <= 256 && N % 8 == 0 && (M == 64 || N % 16 == 0) &&
                  K >= 8 && K % 8 == 0,
              LRX_ERR_INVALID_ARG, "lrx_debug_umma_gemm: unsupported shape M=%d N=%d K=%d", M, N, K);
  SelftestCfg cfg;
  const int32_t* c = cfg_host;
  cfg.a = SelftestOperand{c[0], c[1], c[2], c[3], c[4], c[5]};
  cfg.b = SelftestOperand{c[6], c[7], c[8], c[9], c[10], c[11]};
  const uint32_t bytesA = (c[0] & 2) ? 32768u : selftest_extent(cfg.a, M, K);
  const uint32_t bytesB = (c[6] & 2) ? 32768u : selftest_extent(cfg.b, N, K);
  const size_t smem = (size_t)2 * (bytesA + bytesB);
  LRX_REQUIRE(smem <= 200 * 1024, LRX_ERR_INVALID_ARG, "lrx_debug_umma_gemm: operands exceed shared memory");
  LRX_CUDA_CHECK(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, dump, M, N, K, cfg, split, bytesA, bytesB);
  LRX_LAUNCH_CHECK();
  return LRX_OK;
}

// ------------------------------------------------------------------ bf16x3 variant
// fp32 emulation with bf16 pieces: x = p0 + p1 + p2 (8 mantissa bits each, truncation), products
// p_i * q_j kept for i + j <= 2 (6 MMAs, kind::f16, fp32 accumulation).  Element size 2 bytes:
// a 16-byte chunk holds 8 consecutive k (K-major) or 8 consecutive mn (MN-major); K = 16 per MMA.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

// fp16x2 probe: x = h0 + h1 with h0 = fp16_rn(x), h1 = fp16_rn(x - h0); stored in the same 2-byte slots
__device__ __forceinline__ void fp16_split2(float x, __nv_bfloat16 (&p)[3]) {
  const __half h0 = __float2half_rn(x);
  const __half h1 = __float2half_rn(x - __half2float(h0));
  p[0] = *reinterpret_cast<const __nv_bfloat16*>(&h0);
  p[1] = *reinterpret_cast<const __nv_bfloat16*>(&h1);
  p[2] = __float2bfloat16_rn(0.f);
}
__device__ __forceinline__ void bf16_split3(float x, __nv_bfloat16 (&p)[3]) {
  float r = x;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float h = __uint_as_float(__float_as_uint(r) & 0xFFFF0000u);  // exact bf16 (truncation)
    p[i] = __float2bfloat16_rn(h);
    r -= h;
  }
}

__device__ __forceinline__ uint64_t tc_desc(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

__device__ __forceinline__ uint32_t selftest_off16(const SelftestOperand& o, int mn, int k) {
  return o.mn ? (uint32_t)((mn & 7) * 2 + (mn >> 3) * o.s_mn + (k & 7) * 16 + (k >> 3) * o.s_k)
              : (uint32_t)((mn & 7) * 16 + (mn >> 3) * o.s_mn + (k & 7) * 2 + (k >> 3) * o.s_k);
}

__global__ void __launch_bounds__(128)
umma_selftest_bf16_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ dump,
                          int M, int N, int K, SelftestCfg cfg, int nprod, uint32_t bytesA, uint32_t bytesB,
                          int reps, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool raw_a = cfg.a.mn & 2, raw_b = cfg.b.mn & 2;
  const uint32_t lt_a = (cfg.a.mn >> 4) & 7, lt_b = (cfg.b.mn >> 4) & 7;
  cfg.a.mn &= 1;
  cfg.b.mn &= 1;
  uint8_t* a_p[3] = {smem, smem + bytesA, smem + 2 * bytesA};
  uint8_t* b_p[3] = {smem + 3 * bytesA, smem + 3 * bytesA + bytesB, smem + 3 * bytesA + 2 * bytesB};
  __nv_bfloat16 pc[3];
  const bool fp16_mode = (nprod & 0x10000) != 0;
  if (raw_a) {
    for (uint32_t w = tid; w < bytesA / 2; w += blockDim.x) {
      bf16_split3((float)w, pc);
      for (int i = 0; i < 3; ++i) ((__nv_bfloat16*)a_p[i])[w] = pc[i];
    }
  } else {
    for (int i = tid; i < M * K; i += blockDim.x) {
      if (fp16_mode) fp16_split2(A[i], pc); else bf16_split3(A[i], pc);
      const uint32_t off = selftest_off16(cfg.a, i / K, i % K);
      for (int q = 0; q < 3; ++q) *(__nv_bfloat16*)(a_p[q] + off) = pc[q];
    }
  }
  if (raw_b) {
    for (uint32_t w = tid; w < bytesB / 2; w += blockDim.x) {
      bf16_split3((float)w, pc);
      for (int i = 0; i < 3; ++i) ((__nv_bfloat16*)b_p[i])[w] = pc[i];
    }
  } else {
    for (int i = tid; i < N * K; i += blockDim.x) {
      if (fp16_mode) fp16_split2(B[i], pc); else bf16_split3(B[i], pc);
      const uint32_t off = selftest_off16(cfg.b, i / K, i % K);
      for (int q = 0; q < 3; ++q) *(__nv_bfloat16*)(b_p[q] + off) = pc[q];
    }
  }
  if (tid == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 256);
  umma::fence_async_smem();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tbase = tmem_slot;
  {
    float z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0.f;
    for (int c = 0; c < 256; c += 16) umma::tmem_st16(umma::tmem_addr(tbase, warp * 32, c), z);
    umma::tmem_st_wait();
  }
  umma::tc_fence_before();
  __syncthreads();

  if (umma::warp_idx_sync() == 0) {
    umma::tc_fence_after();
    const bool leader = umma::elect_one();
    uint32_t idesc = umma::idesc_bf16(M, N, cfg.a.mn, cfg.b.mn);
    if (fp16_mode) idesc &= ~((7u << 7) | (7u << 10));  // a_format = b_format = F16
    const uint32_t ahi = ((uint32_t)cfg.a.sbo >> 4) | (1u << 14) | (lt_a << 29);
    const uint32_t bhi = ((uint32_t)cfg.b.sbo >> 4) | (1u << 14) | (lt_b << 29);
    const uint32_t alo0 = ((umma::smem_u32(a_p[0]) >> 4) & 0x3FFFu) | (((uint32_t)cfg.a.lbo >> 4) << 16);
    const uint32_t blo0 = ((umma::smem_u32(b_p[0]) >> 4) & 0x3FFFu) | (((uint32_t)cfg.b.lbo >> 4) << 16);
    const uint32_t aps = bytesA >> 4, bps = bytesB >> 4, aks = (uint32_t)cfg.a.kstep >> 4, bks = (uint32_t)cfg.b.kstep >> 4;
    uint32_t acc = 0;
    const int dsplit = ((nprod >> 8) & 0xFF) ? ((nprod >> 8) & 0xFF) : 1;  // timing: rotate over independent accumulators
    nprod &= 0xFF;
    uint32_t dn = 0;
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
      uint32_t alo = alo0, blo = blo0;
      for (int ks = 0; ks < K / 16; ++ks) {
        // smallest products first: i + j = 2, then 1, then 0
        for (int sum = (nprod >= 6 ? 2 : (nprod >= 3 ? 1 : 0)); sum >= 0; --sum)
          for (int i = 0; i <= sum; ++i) {
            const int j = sum - i;
            if (leader)
              umma::mma_f16(tbase + dn * N, tc_desc(alo + i * aps, ahi), tc_desc(blo + j * bps, bhi), idesc, acc);
            dn = (dn + 1 == (uint32_t)dsplit) ? 0u : dn + 1;
            acc = (reps > 1) ? 0u : 1u;  // timing mode overwrites (keeps values finite)
          }
        alo += aks;
        blo += bks;
      }
    }
    const long long t1 = clock64();
    if (leader) umma::mma_commit(&bar);
    __syncwarp();
    umma::mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (leader && cycles_out) { cycles_out[0] = t1 - t0; cycles_out[1] = t2 - t0; }
  }
  umma::mbar_wait(&bar, 0);
  umma::tc_fence_after();
  for (int c = 0; c < N; c += 8) {
    float v[8];
    umma::tmem_ld8(umma::tmem_addr(tbase, warp * 32, c), v);
    umma::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) dump[(size_t)tid * N + c + i] = v[i];
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 256);
}

static uint32_t selftest_extent16(SelftestOperand o, int MN, int K) {
  o.mn &= 1;
  uint32_t mx = 0;
  for (int mn = 0; mn < MN; ++mn)
    for (int k = 0; k < K; ++k) {
      uint32_t off = o.mn ? (uint32_t)((mn & 7) * 2 + (mn >> 3) * o.s_mn + (k & 7) * 16 + (k >> 3) * o.s_k)
                          : (uint32_t)((mn & 7) * 16 + (mn >> 3) * o.s_mn + (k & 7) * 2 + (k >> 3) * o.s_k);
      if (off + 2 > mx) mx = off + 2;
    }
  return (mx + 1023u) / 1024u * 1024u;
}

// Same contract as lrx_debug_umma_gemm for the bf16x3 emulation (K multiple of 16; nprod = 1, 3 or 6
// products kept).
extern "C" int lrx_debug_umma_gemm_bf16(const float* A, const float* B, float* dump, int32_t M, int32_t N,
                                        int32_t K, const int32_t* cfg_host, int32_t nprod, lrx_stream_t stream) {
  LRX_REQUIRE(A && B && dump && cfg_host, LRX_ERR_INVALID_ARG, "lrx_debug_umma_gemm_bf16: NULL pointer");
  LRX_REQUIRE((M == 64 || M == 128) && N >= 8 && N <= 256 && N % 8 == 0 && (M == 64 || N % 16 == 0) &&
                  K >= 16 && K % 16 == 0,
              LRX_ERR_INVALID_ARG, "lrx_debug_umma_gemm_bf16: unsupported shape M=%d N=%d K=%d", M, N, K);
  SelftestCfg cfg;
  const int32_t* c = cfg_host;
  cfg.a = SelftestOperand{c[0], c[1], c[2], c[3], c[4], c[5]};
  cfg.b = SelftestOperand{c[6], c[7], c[8], c[9], c[10], c[11]};
  const uint32_t bytesA = (c[0] & 2) ? 16384u : selftest_extent16(cfg.a, M, K);
  const uint32_t bytesB = (c[6] & 2) ? 16384u : selftest_extent16(cfg.b, N, K);
  const size_t smem = (size_t)3 * (bytesA + bytesB);
  LRX_REQUIRE(smem <= 200 * 1024, LRX_ERR_INVALID_ARG, "lrx_debug_umma_gemm_bf16: operands exceed shared memory");
  LRX_CUDA_CHECK(cudaFuncSetAttribute(umma_selftest_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_bf16_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, dump, M, N, K, cfg, nprod & 0x100FF, bytesA,
                                                                      bytesB, 1, nullptr);
  LRX_LAUNCH_CHECK();
  return LRX_OK;
}

// Timing probe: the same MMA sequence repeated `reps` times by one thread; cycles[0] = issue time,
// cycles[1] = issue + completion (SM clocks).  Used to size the fused kernels (tools/umma_time.py).
extern "C" int lrx_debug_umma_time_bf16(const float* A, const float* B, float* dump, int32_t M, int32_t N, int32_t K,
                                        const int32_t* cfg_host, int32_t nprod, int32_t reps, long long* cycles,
                                        lrx_stream_t stream) {
  LRX_REQUIRE(A && B && dump && cfg_host && cycles && reps >= 1, LRX_ERR_INVALID_ARG, "lrx_debug_umma_time_bf16: bad argument");
  SelftestCfg cfg;
  const int32_t* c = cfg_host;
  cfg.a = SelftestOperand{c[0], c[1], c[2], c[3], c[4], c[5]};
  cfg.b = SelftestOperand{c[6], c[7], c[8], c[9], c[10], c[11]};
  const uint32_t bytesA = selftest_extent16(cfg.a, M, K), bytesB = selftest_extent16(cfg.b, N, K);
  const size_t smem = (size_t)3 * (bytesA + bytesB);
  LRX_REQUIRE(smem <= 200 * 1024, LRX_ERR_INVALID_ARG, "lrx_debug_umma_time_bf16: operands exceed shared memory");
  LRX_CUDA_CHECK(cudaFuncSetAttribute(umma_selftest_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_bf16_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, dump, M, N, K, cfg, nprod, bytesA, bytesB,
                                                                      reps, cycles);
  LRX_LAUNCH_CHECK();
  return LRX_OK;
}

// ------------------------------------------------------------------ MMA dispatch-rate probe
// 256 back-to-back tcgen05.mma with CONSTANT descriptors (no per-instruction arithmetic) rotating
// over DS independent accumulators: the tensor pipe's own cost per instruction for small shapes.
template <int DS>
__global__ void __launch_bounds__(128) umma_rate_kernel(int M, int N, int a_mn, int b_mn, long long* cycles_out, int a_tmem = 0,
                                                        int issuers = 1) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_slot;
  const int warp = umma::warp_idx_sync();
  for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3F803F80u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) umma::mbar_init(&bar[i], 1);
    umma::fence_mbar_init();
  }
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 512);
  umma::fence_async_smem();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tbase = tmem_slot;
  if (warp < issuers) {  // `issuers` warps issue concurrently, each into its own accumulators and its own barrier
    const bool leader = umma::elect_one();
    // K-major view: LBO = rows * 16, SBO = 128; MN-major view of the same canonical bytes: LBO = 128, SBO = rows * 16
    const uint32_t idesc = umma::idesc_bf16(M, N, a_mn, b_mn);
    const uint64_t ad = a_mn ? umma::smem_desc(umma::smem_u32(smem), 128u, (uint32_t)M * 16u)
                             : umma::smem_desc(umma::smem_u32(smem), (uint32_t)M * 16u, 128u);
    const uint64_t bd = b_mn ? umma::smem_desc(umma::smem_u32(smem) + 8192u, 128u, (uint32_t)N * 16u)
                             : umma::smem_desc(umma::smem_u32(smem) + 8192u, (uint32_t)N * 16u, 128u);
    const uint32_t dbase = tbase + (uint32_t)warp * 64u * (uint32_t)DS;
    const long long t0 = clock64();
    if (a_tmem) {  // TS form: the A operand (bf16 pairs in 32-bit columns) from TMEM columns 448.., B from shared memory
#pragma unroll 1
      for (int rep = 0; rep < 16; ++rep) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (leader) umma::mma_f16_ts(dbase + (uint32_t)((i % DS) * 64), tbase + 448u, bd, idesc, 1u);
      }
    } else {
#pragma unroll 1
      for (int rep = 0; rep < 16; ++rep) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (leader) umma::mma_f16(dbase + (uint32_t)((i % DS) * 64), ad, bd, idesc, 1u);
      }
    }
    const long long t1 = clock64();
    if (leader) umma::mma_commit(&bar[warp]);
    __syncwarp();
    umma::mbar_wait(&bar[warp], 0);
    const long long t2 = clock64();
    if (leader) { cycles_out[2 * warp] = t1 - t0; cycles_out[2 * warp + 1] = t2 - t0; }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 512);
}

extern "C" int lrx_debug_umma_rate(int32_t M, int32_t N, int32_t dsplit, long long* cycles, lrx_stream_t stream) {
  LRX_REQUIRE(cycles && (M == 64 || M == 128) && N >= 8 && N <= 64 && N % 8 == 0 && (M == 64 || N % 16 == 0),
              LRX_ERR_INVALID_ARG, "lrx_debug_umma_rate: bad argument");
  const size_t smem = 16384;
  cudaStream_t s = (cudaStream_t)stream;
  const int a_mn = (dsplit >> 8) & 1, b_mn = (dsplit >> 9) & 1;  // operand majorness rides in the upper bits
  const int a_tmem = (dsplit >> 10) & 1;
  const int issuers = 1 + ((dsplit >> 11) & 3);  // `cycles` must hold 2 * issuers values
  dsplit &= 0xff;
  if (dsplit == 1) umma_rate_kernel<1><<<1, 128, smem, s>>>(M, N, a_mn, b_mn, cycles, a_tmem, issuers);
  else if (dsplit == 2) umma_rate_kernel<2><<<1, 128, smem, s>>>(M, N, a_mn, b_mn, cycles, a_tmem, issuers);
  else if (dsplit == 4) umma_rate_kernel<4><<<1, 128, smem, s>>>(M, N, a_mn, b_mn, cycles, a_tmem, issuers);
  else umma_rate_kernel<8><<<1, 128, smem, s>>>(M, N, a_mn, b_mn, cycles, a_tmem, issuers);
  LRX_LAUNCH_CHECK();
  return LRX_OK;
}

// ------------------------------------------------------------------ tcgen05.ld latency under MMA load
// Warp 0 enqueues n_mma MMAs (M = 64, N = 64, accumulating into TMEM columns 0-63) and commits; then warps 0-3
// time eight tcgen05.ld + wait::ld of ANOTHER TMEM region (columns 256+).  cycles[0..7] = latency of the k-th load
// of warp 1, cycles[8] = cycles until the MMA batch's mbarrier completes (from the first load).
__global__ void __launch_bounds__(128) tmem_ld_under_mma_kernel(int n_mma_signed, long long* cycles_out) {
  const bool acc_first = (n_mma_signed >= 10000);  // +10000: the first MMA accumulates too (onto uninitialised TMEM)
  if (acc_first) n_mma_signed -= 10000;
  const int n_mma = n_mma_signed < 0 ? -n_mma_signed : n_mma_signed;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = umma::warp_idx_sync();
  for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3F803F80u;
  if (threadIdx.x == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 512);
  umma::fence_async_smem();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tbase = tmem_slot;
  if (warp == 0) {
    const bool leader = umma::elect_one();
    const uint32_t idesc = umma::idesc_bf16(64, 64, 0, 0);
    const uint64_t ad = umma::smem_desc(umma::smem_u32(smem), 64u * 16u, 128u);
    const uint64_t bd = umma::smem_desc(umma::smem_u32(smem) + 8192u, 64u * 16u, 128u);
#pragma unroll 1
    for (int i = 0; i < n_mma; ++i)
      if (leader) umma::mma_f16(tbase, ad, bd, idesc, (i > 0 || acc_first) ? 1u : 0u);
    if (leader) umma::mma_commit(&bar);
    __syncwarp();
  }
  __syncthreads();
  // n_mma < 0: load the ACCUMULATOR itself (columns 0-63) without waiting for the commit: cycles[9] = the value read
  // (every MMA adds 16 to each element), to see whether the load is ordered behind the queued MMAs' writes
  const bool own = n_mma_signed < 0;
  const uint32_t tw = umma::tmem_addr(tbase, warp * 32, own ? 0 : 256);
  long long lat[8];
  float first = 0.f;
  const long long t_start = clock64();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const long long t0 = clock64();
    float v[8];
    umma::tmem_ld_frag16(tw + (own ? 16 * (k & 3) : 16 * k), v);
    umma::tmem_ld_wait();
    if (k == 0) first = v[0];
    lat[k] = clock64() - t0 + (long long)(v[0] == 12345.f);
  }
  if (n_mma > 0) umma::mbar_wait(&bar, 0);
  const long long t_done = clock64() - t_start;
  if (threadIdx.x == 32) {
    for (int k = 0; k < 8; ++k) cycles_out[k] = lat[k];
    cycles_out[8] = t_done;
    cycles_out[9] = (long long)first;
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 512);
}

extern "C" int lrx_debug_tmem_ld_under_mma(int32_t n_mma, long long* cycles, lrx_stream_t stream) {
  LRX_REQUIRE(cycles && n_mma >= -4096 && n_mma <= 14096, LRX_ERR_INVALID_ARG, "lrx_debug_tmem_ld_under_mma: bad argument");
  tmem_ld_under_mma_kernel<<<1, 128, 16384, (cudaStream_t)stream>>>(n_mma, cycles);
  LRX_LAUNCH_CHECK();
  return LRX_OK;
}

// ------------------------------------------------------------------ tcgen05.ld fragment-shape probe
// TMEM[lane][col] = lane * 100 + col (written 32x32b), read back with the 16x256b.x2 shape: dump[t][r]
// = register r of thread t, which decodes the (lane, column) each thread receives.
// variant 1: the STORE shape 16x128b.x1 (two registers per thread, value = 2 * thread + register), read back 32x32b:
// dump[lane][col] = which (thread, register) wrote TMEM cell (lane, col), cols 0..3.
__global__ void __launch_bounds__(128) tmem_st_shape_probe_kernel(float* __restrict__ dump) {
  __shared__ uint32_t tmem_slot;
  const int warp = umma::warp_idx_sync(), lane = threadIdx.x & 31;
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 32);
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tb = tmem_slot;
  float z[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) z[c] = -1.f;
  umma::tmem_st16(umma::tmem_addr(tb, warp * 32, 0), z);
  umma::tmem_st_wait();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t r0 = __float_as_uint((float)(2 * lane)), r1 = __float_as_uint((float)(2 * lane + 1));
  asm volatile("tcgen05.st.sync.aligned.16x128b.x1.b32 [%0], {%1, %2};" ::"r"(umma::tmem_addr(tb, warp * 32, 0)), "r"(r0), "r"(r1)
               : "memory");
  umma::tmem_st_wait();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  float v[8];
  umma::tmem_ld8(umma::tmem_addr(tb, warp * 32, 0), v);
  umma::tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 8; ++i) dump[threadIdx.x * 8 + i] = v[i];
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tb, 32);
}

__global__ void __launch_bounds__(128) tmem_shape_probe_kernel(float* __restrict__ dump) {
  __shared__ uint32_t tmem_slot;
  const int warp = umma::warp_idx_sync(), lane = threadIdx.x & 31;
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 32);
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tb = tmem_slot;
  float v[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) v[c] = (float)((warp * 32 + lane) * 100 + c);
  umma::tmem_st16(umma::tmem_addr(tb, warp * 32, 0), v);
  umma::tmem_st_wait();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(umma::tmem_addr(tb, warp * 32, 0)));
  umma::tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 8; ++i) dump[threadIdx.x * 8 + i] = __uint_as_float(r[i]);
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tb, 32);
}
extern "C" int lrx_debug_tmem_shape_probe(float* dump, lrx_stream_t stream) {
  LRX_REQUIRE(dump, LRX_ERR_INVALID_ARG, "lrx_debug_tmem_shape_probe: NULL pointer");
  if (getenv("LRX_PROBE_ST_SHAPE")) {
    tmem_st_shape_probe_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(dump);
    LRX_LAUNCH_CHECK();
    return LRX_OK;
  }
  tmem_shape_probe_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(dump);
  LRX_LAUNCH_CHECK();
  return LRX_OK;
}

// ------------------------------------------------------------------ A-from-TMEM / accumulator-lane probes
// mode 0: D = A_tmem * I  with A written to TMEM by tcgen05.st 32x32b as bf16 pairs: word(lane L, col c) =
//         (bf16(f(L, 2c)) | bf16(f(L, 2c+1)) << 16), f = lane id (variant 0) or k index (variant 1); B = identity
//         [16][16] in shared memory.  The accumulator dump shows which TMEM lane / half-word feeds row m, column k.
// mode 1: SS-mode M = 64 MMA (A = row id, B = identity) with the accumulator address at TMEM lane 16: does an
//         M = 64 accumulator live in the upper half of each 32-lane quadrant?
__global__ void __launch_bounds__(128) tmem_operand_probe_kernel(float* __restrict__ dump, int mode, int variant, int M) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = umma::warp_idx_sync(), lane = threadIdx.x & 31, tid = threadIdx.x;
  __nv_bfloat16* Bm = reinterpret_cast<__nv_bfloat16*>(smem);            // identity [16][16], canonical K-major
  __nv_bfloat16* Am = reinterpret_cast<__nv_bfloat16*>(smem + 1024);     // [M][16] row id, canonical K-major (mode 1)
  for (int i = tid; i < 16 * 16; i += 128) {
    const int n = i / 16, k = i % 16;
    Bm[((k >> 3) * (16 * 16) + n * 16 + (k & 7) * 2) / 2] = __float2bfloat16_rn(n == k ? 1.f : 0.f);
  }
  for (int i = tid; i < 128 * 16; i += 128) {
    const int m = i / 16, k = i % 16;
    Am[((k >> 3) * (128 * 16) + m * 16 + (k & 7) * 2) / 2] = __float2bfloat16_rn(k == 0 ? (float)m : 0.f);
  }
  if (tid == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 64);
  umma::fence_async_smem();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tb = tmem_slot;
  {
    float z[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) z[c] = 0.f;
    for (int c = 0; c < 64; c += 16) umma::tmem_st16(umma::tmem_addr(tb, warp * 32, c), z);
    // A operand at columns 32..39: 8 packed words per lane
    float w[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float lo = variant ? (float)(2 * c) : (float)tid, hi = variant ? (float)(2 * c + 1) : (float)tid;
      const uint32_t word = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(lo)) |
                            ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(hi)) << 16);
      w[c] = __uint_as_float(word);
    }
    umma::tmem_st16(umma::tmem_addr(tb, warp * 32, 32), w);
    umma::tmem_st_wait();
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    umma::tc_fence_after();
    const bool leader = umma::elect_one();
    const uint32_t idesc = umma::idesc_bf16(M, 16, 0, 0);
    const uint64_t bd = umma::smem_desc(umma::smem_u32(Bm), 16u * 16u, 128u);
    if (leader) {
      if (mode == 0) {
        const uint32_t a_t = tb + 32;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tb),
            "r"(a_t), "l"(bd), "r"(idesc), "r"(0u)
            : "memory");
      } else {
        const uint64_t ad = umma::smem_desc(umma::smem_u32(Am), 128u * 16u, 128u);
        umma::mma_f16(tb + (16u << 16), ad, bd, idesc, 0u);
      }
      umma::mma_commit(&bar);
    }
    __syncwarp();
  }
  umma::mbar_wait(&bar, 0);
  umma::tc_fence_after();
  float v[16];
  umma::tmem_ld16(umma::tmem_addr(tb, warp * 32, 0), v);
  umma::tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) dump[tid * 16 + i] = v[i];
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tb, 64);
}
extern "C" int lrx_debug_tmem_operand_probe(float* dump, int32_t mode, int32_t variant, int32_t M, lrx_stream_t stream) {
  LRX_REQUIRE(dump && (M == 64 || M == 128), LRX_ERR_INVALID_ARG, "lrx_debug_tmem_operand_probe: bad argument");
  tmem_operand_probe_kernel<<<1, 128, 8192, (cudaStream_t)stream>>>(dump, mode, variant, M);
  LRX_LAUNCH_CHECK();
  return LRX_OK;
}
