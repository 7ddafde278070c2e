// Generic-width tcgen05 GEMM for the layer-by-layer PPO update (any MLPActorCriticPolicy widths; BASELINE.json
// configs[2]: Pendulum, 2-layer 256-wide MLPs).  Drop-in for sgemm_kernel of lrx_update.cu: same GemmArgs, same
// epilogue (bias, ReLU, ReLU mask, accumulate, ones row, redirected bias-gradient column, split-K).
//
//     C[M x N] (+)= op(A)[M x K] * op(B)[K x N]      fp32 in HBM, two-piece bf16 emulation on the tensor cores
//
// One CTA = one 128 x (<= 256) accumulator in TMEM.  The inputs are fp32 in global memory in either orientation; a K
// chunk of 64 of both operands is read with 32-byte accesses along the contiguous axis, split into bf16 hi / lo
// (lrx_tc2.cuh) and written into the canonical no-swizzle layout, which serves as a K-major operand when K is the
// contiguous axis and as an MN-major operand when the M/N index is -- so all four orientations of sgemm_kernel<AT, BT>
// feed tcgen05.mma without a transpose.  SWAP decides which index of C rides on the 128 TMEM lanes:
//   SWAP = 1 (forward / dX: N = minibatch rows, M = layer width): lanes = j, columns = i -> C[i][j] is written with
//            consecutive lanes on consecutive addresses (coalesced);
//   SWAP = 0 (weight gradients: M = out, N = in + 1, K = rows, split-K): lanes = i, columns = j.
// A CTA is not software pipelined: stage -> MMA -> commit -> next chunk; two CTAs share an SM (96 KB of shared memory,
// 256 TMEM columns each) so that one's staging runs under the other's MMAs.
#include "lrx_common.cuh"
#include "lrx_gemm.cuh"
#include "lrx_tc2.cuh"

namespace {

constexpr int KC = 64;          // K chunk
constexpr int UM = 128;         // UMMA M (TMEM lanes)
constexpr int UN_MAX = 256;     // UMMA N per accumulator
constexpr int THREADS = 256;
// shared memory: operand X (UMMA A, 128 mn) and operand Y (UMMA B, <= 256 mn), each hi | lo
constexpr uint32_t X_PIECE = UM * KC * 2, Y_PIECE = UN_MAX * KC * 2;
constexpr uint32_t OFF_X = 0, OFF_Y = OFF_X + 2 * X_PIECE, OFF_MISC = OFF_Y + 2 * Y_PIECE;
constexpr uint32_t SMEM_BYTES = OFF_MISC + 64;

// Source of one UMMA operand: element (mn, k) of op(.) lives at base[mn * s_mn + k * s_k]; exactly one stride is 1.
struct Src {
  const float* base;
  long long s_mn, s_k;
  int mn_total;
};

// Stage rows [mn0, mn0 + mn_tile) x [k0, k0 + KC) (zero filled outside [0, mn_total) x [k0, kend)) into `dst`.
//   k contiguous : matrix r = mn, c = k   (R = mn_tile rows)  -> K-major operand
//   mn contiguous: matrix r = k,  c = mn  (R = KC rows)       -> MN-major operand
template <bool MN_CONTIG>
__device__ __forceinline__ void stage(const Src& s, int mn0, int mn_tile, int k0, int kend, const tc2::Mat& dst, int tid) {
  const int R = MN_CONTIG ? KC : mn_tile;
  const int chunks = (MN_CONTIG ? mn_tile : KC) >> 3;
  const int total = R * chunks;
  constexpr int U = 4;  // 16-byte-chunk pairs in flight per thread: the loads of U chunks are issued before any is converted
  for (int e0 = tid; e0 < total; e0 += U * THREADS) {
    float v[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = e0 + u * THREADS;
      const int r = e % R, ch = e / R;
      // (mn, k) of the first of the 8 contiguous elements; `lim` = valid elements along the contiguous axis
      const int mn = MN_CONTIG ? mn0 + ch * 8 : mn0 + r, k = MN_CONTIG ? k0 + r : k0 + ch * 8;
      const bool row_ok = e < total && (MN_CONTIG ? k < kend : mn < s.mn_total);
      const int lim = row_ok ? (MN_CONTIG ? s.mn_total - mn : kend - k) : 0;
      const float* src = s.base + (MN_CONTIG ? (long long)k * s.s_k + mn : (long long)mn * s.s_mn + k);
      if (lim >= 8 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
        v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w; v[u][4] = b.x; v[u][5] = b.y; v[u][6] = b.z; v[u][7] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[u][j] = j < lim ? __ldg(src + j) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = e0 + u * THREADS;
      if (e < total) tc2::store8(dst, e / R, e % R, v[u]);
    }
  }
}

// X_MN / Y_MN: the UMMA A / B operand's source has the MN index contiguous (-> MN-major descriptor).
template <bool SWAP, bool X_MN, bool Y_MN>
__global__ void __launch_bounds__(THREADS, 2) tc_gemm_kernel(const GemmArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = umma::warp_idx_sync(), lane = tid & 31;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_MISC);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + OFF_MISC + 16);

  // tile of C: x -> the index on the TMEM lanes (128 per CTA), y -> the index on the TMEM columns (<= 256 per CTA)
  const int x_total = SWAP ? g.N : g.M, y_total = SWAP ? g.M : g.N;
  const int x0 = blockIdx.x * UM;
  const int y0 = blockIdx.y * UN_MAX;
  const int y_cnt = min(UN_MAX, y_total - y0);
  const int un = (y_cnt + 15) & ~15;  // UMMA N: multiple of 16 for M = 128
  const int kbeg = blockIdx.z * g.k_chunk;
  const int kend = min(g.K, kbeg + g.k_chunk);

  // op(A)[i][k]: AT ? A[k * lda + i] : A[i * lda + k];   op(B)[k][j]: BT ? B[j * ldb + k] : B[k * ldb + j]
  // X carries the lane index (SWAP: j from B, else i from A), Y the column index.
  Src sx, sy;
  if (SWAP) {
    sx = Src{g.B, X_MN ? 1 : g.ldb, X_MN ? g.ldb : 1, g.N};  // mn = j
    sy = Src{g.A, Y_MN ? 1 : g.lda, Y_MN ? g.lda : 1, g.M};  // mn = i
  } else {
    sx = Src{g.A, X_MN ? 1 : g.lda, X_MN ? g.lda : 1, g.M};
    sy = Src{g.B, Y_MN ? 1 : g.ldb, Y_MN ? g.ldb : 1, g.N};
  }
  const tc2::Mat X{smem + OFF_X, (uint32_t)(X_MN ? KC : UM), X_PIECE};
  const tc2::Mat Y{smem + OFF_Y, (uint32_t)(Y_MN ? KC : un), Y_PIECE};

  if (tid == 0) {
    umma::mbar_init(bar, 1);
    umma::fence_mbar_init();
  }
  if (warp == 0) umma::tmem_alloc(tslot, UN_MAX);
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tb = *tslot;

  uint32_t phase = 0, acc = 0;
  for (int k0 = kbeg; k0 < kend; k0 += KC) {
    stage<X_MN>(sx, x0, UM, k0, kend, X, tid);
    stage<Y_MN>(sy, y0, un, k0, kend, Y, tid);
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      umma::tc_fence_after();
      const bool leader = umma::elect_one();
      const int ksteps = (min(KC, kend - k0) + 15) >> 4;  // zero-filled beyond kend
      tc2::gemm<X_MN, Y_MN>(leader, tb, X, Y, UM, un, un, ksteps, acc);
      if (leader) umma::mma_commit(bar);
      __syncwarp();
    }
    acc = 1u;
    umma::mbar_wait(bar, phase);  // the MMAs have read the staging buffers: free for the next chunk
    phase ^= 1u;
    umma::tc_fence_after();
  }

  // ---- epilogue: warp w reads TMEM quadrant w & 3 (lanes = x), column half w >> 2
  const int q = warp & 3, half = warp >> 2;
  const int x = x0 + q * 32 + lane;
  const uint32_t tw = umma::tmem_addr(tb, q * 32, 0);
  const int i_of = SWAP ? 0 : x, j_of = SWAP ? x : 0;  // the fixed index of this thread
  float* C = g.C + (size_t)blockIdx.z * g.c_split_stride;
  float* C2 = g.C2 ? g.C2 + (size_t)blockIdx.z * g.c_split_stride : nullptr;
  const bool any_k = kend > kbeg;
  for (int c0 = half * 16; c0 < un; c0 += 32) {
    float v[16];
    if (any_k) {
      umma::tmem_ld16(tw + c0, v);
      umma::tmem_ld_wait();
    } else {
#pragma unroll
      for (int t = 0; t < 16; ++t) v[t] = 0.f;
    }
    if (x < x_total) {
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const int y = y0 + c0 + t;
        if (y < y_total) {
          const int i = SWAP ? y : i_of, j = SWAP ? j_of : y;
          float r = v[t] + (g.bias ? __ldg(g.bias + i) : 0.f);
          if (g.relu) r = fmaxf(r, 0.f);
          if (g.mask) r = (g.mask[(size_t)i * g.ldm + j] > 0.f) ? r : 0.f;
          float* dst = (C2 && j == g.n_main) ? (C2 + i) : (C + (size_t)i * g.ldc + j);
          if (g.accumulate) r += *dst;
          *dst = r;
        }
      }
    }
  }
  if (g.ones_row && blockIdx.z == 0 && (SWAP ? blockIdx.y == 0 : blockIdx.x == 0)) {
    // C[M][j] = 1 for the j of this CTA's tile
    const int jn = SWAP ? UM : y_cnt, jb = SWAP ? x0 : y0;
    for (int t = tid; t < jn; t += THREADS)
      if (jb + t < g.N) g.C[(size_t)g.M * g.ldc + jb + t] = 1.0f;
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tb, UN_MAX);
}

template <bool SWAP, bool X_MN, bool Y_MN>
int launch(const GemmArgs& g, int splits, cudaStream_t s) {
  static bool attr_set[64];
  LRX_CUDA_CHECK(lrx_set_max_smem(tc_gemm_kernel<SWAP, X_MN, Y_MN>, (int)SMEM_BYTES, attr_set));
  const int x_total = SWAP ? g.N : g.M, y_total = SWAP ? g.M : g.N;
  dim3 grid((x_total + UM - 1) / UM, (y_total + UN_MAX - 1) / UN_MAX, splits);
  tc_gemm_kernel<SWAP, X_MN, Y_MN><<<grid, THREADS, SMEM_BYTES, s>>>(g);
  return LRX_OK;
}

}  // namespace

// Returns LRX_OK and sets *handled when the shape is worth the tensor cores (both the contraction and the layer width at
// least 16; the d <= 8 first layer and the scalar heads stay on the FMA kernel).
int lrx_launch_gemm_tc(const GemmArgs& g, bool at, bool bt, int splits, cudaStream_t s, bool* handled) {
  *handled = false;
  if (lrx_fused_enabled() < 1) return LRX_OK;
  const bool rows_on_n = g.N >= g.M;  // forward / dX: the minibatch rows are N; weight gradients: the rows are K
  // narrow layers (the d <= 8 first layer, scalar heads) ride along zero padded to the 16-wide MMA minimum: their cost is
  // the HBM pass over the wide operand, which the FMA kernel made at a sixth of the bandwidth.  Tiny problems stay on it.
  const long long big = (long long)(g.M > g.N ? g.M : g.N) * (g.K > 16 ? g.K : 16);
  if (g.M < 1 || g.N < 1 || g.K < 1 || big < 128 * 64) return LRX_OK;
  int rc;
  if (rows_on_n) {
    // SWAP: X = op(B) (mn = j): contiguous in j unless BT;  Y = op(A) (mn = i): contiguous in i iff AT
    if (!bt && !at) rc = launch<true, true, false>(g, splits, s);
    else if (!bt && at) rc = launch<true, true, true>(g, splits, s);
    else if (bt && !at) rc = launch<true, false, false>(g, splits, s);
    else rc = launch<true, false, true>(g, splits, s);
  } else {
    // X = op(A) (mn = i): contiguous in i iff AT;  Y = op(B) (mn = j): contiguous in j unless BT
    if (!at && bt) rc = launch<false, false, false>(g, splits, s);
    else if (!at && !bt) rc = launch<false, false, true>(g, splits, s);
    else if (at && bt) rc = launch<false, true, false>(g, splits, s);
    else rc = launch<false, true, true>(g, splits, s);
  }
  if (rc) return rc;
  LRX_LAUNCH_CHECK();
  *handled = true;
  return LRX_OK;
}
