// Fused PPO minibatch gradient for the default MLPActorCriticPolicy widths, third generation: TWO 64-row half tiles in
// flight per SM ("ping-pong"), two-piece bf16 arithmetic (lrx_tc2.cuh), one shared issuer warp.
//
// Reference: algorithm/ppo.py:396-469 (ppo_loss + filter_value_and_grad) over the gathered minibatch of
// buffer/base_buffer.py:90-103 -- the same contract as lrx_update_tc128.cu / lrx_update_fused.cu, which stay selectable
// (lrx_debug_set_fused 2 / 1) and are checked against this kernel in the tests.
//
// Why.  The 128-row kernel is bound by a serial chain of twelve steps per tile (critical MMA group -> commit -> tcgen05.ld
// -> fp32 epilogue -> bf16 split -> st.shared -> barrier): tools/tc128_prof.py gives 25.5 k cycles per tile with the
// tensor pipe 19 % and the issue slots 42 % busy -- every step leaves the SM waiting for ~0.9 k cycles of MMA + commit
// latency, and three steps (encoder layer 0, value loss, policy loss) have no tensor work at all.  A second 128-row tile
// does not fit in shared memory (156 KB of activations per tile next to 49 KB of weights), but two 64-row tiles do: the
// same bytes, split in two independent halves that walk the chain out of step.  While one half's eight warps run an
// epilogue, the other half's GEMMs occupy the tensor pipe; the weights (shared memory) and the weight-gradient
// accumulators (TMEM, M = 64) are shared by both halves, the issuer warp serves the two request streams in arrival order.
//
// Half tile = UMMA M = 64: row j of an accumulator sits in TMEM lane 32 * (j / 16) + j % 16.  Epilogue warp (q, ch) of a
// half reads its quadrant's 16 rows with the 16x256b load shape (all 32 lanes hold data: thread t has rows t/4 and
// t/4 + 8, column pairs 2 (t % 4) of every 8-column chunk) and 32 of a 64-wide layer's columns (ch = 0 / 1).
#include <stdlib.h>

#include "lrx_common.cuh"
#include "lrx_policy.cuh"
#include "lrx_tc2.cuh"

namespace {

constexpr int HROWS = 64;              // rows per half tile = UMMA M
constexpr int HWARPS = 8;              // epilogue warps per half
constexpr int HTHREADS = HWARPS * 32;  // 256
constexpr int WORKERS = 2 * HTHREADS;  // 512
constexpr int THREADS = WORKERS + 32;  // + warp 16: issues every tcgen05.mma of both halves

// ---- shared memory map (bytes).  Weights as in lrx_update_tc128.cu; a c-group (8 columns) of a 64-row matrix is 1 KB.
constexpr uint32_t CG = HROWS * 16;
constexpr uint32_t OFF_W1 = 0;                       // [64][64]  hi 8192 | lo 8192
constexpr uint32_t OFF_W2 = OFF_W1 + 16384;          // [16][64]  hi 2048 | lo 2048
constexpr uint32_t OFF_WV0 = OFF_W2 + 4096;          // [64][16]
constexpr uint32_t OFF_WV1 = OFF_WV0 + 4096;         // [64][64]
constexpr uint32_t OFF_WA0 = OFF_WV1 + 16384;        // [64][16]
constexpr uint32_t OFF_WA1 = OFF_WA0 + 4096;         // [16][64]
constexpr uint32_t OFF_HALF = OFF_WA1 + 4096;        // two per-half regions follow
// per-half region (offsets from the half's base)
constexpr uint32_t H_XOBS = 0;                       // [64][8 obs | ones, 0 x 7]: hi 2 c-groups, lo 1
constexpr uint32_t H_XH1 = H_XOBS + 3 * CG;          // [64][64 | ones]: hi 9, lo 8
constexpr uint32_t H_XH2 = H_XH1 + 17 * CG;          // [64][64]: hi 8, lo 8 (also d(loss)/d(h1) at the end of a tile)
constexpr uint32_t H_XF = H_XH2 + 16 * CG;           // [64][16 | ones]: hi 3, lo 2
constexpr uint32_t H_XB1 = H_XF + 5 * CG;            // [64][64 | ones]: head layer-1 activations (value, then action)
constexpr uint32_t H_DY64 = H_XB1 + 17 * CG;         // [64][64]: current upstream gradient
constexpr uint32_t H_DY16 = H_DY64 + 16 * CG;        // [64][16]
constexpr int R_RET = 0, R_VAL = 1, R_ACT = 2, R_LOGP = 3, R_ADV = 4, R_PLANES = 5;
constexpr uint32_t H_OBS32 = H_DY16 + 4 * CG;        // fp32 [64][8] gathered observations (single-buffered)
constexpr uint32_t H_PLANES = H_OBS32 + HROWS * 8 * 4;
constexpr uint32_t PLANES_BYTES = R_PLANES * HROWS * 4;  // double-buffered
constexpr uint32_t H_VALID = H_PLANES + 2 * PLANES_BYTES;
constexpr uint32_t H_IDX = H_VALID + 2 * HROWS;      // minibatch indices of the tile after the one being gathered
constexpr uint32_t H_VPART = H_IDX + HROWS * 4;      // [2 column halves][64 rows] partial value dot products
constexpr uint32_t H_BYTES = H_VPART + 2 * HROWS * 4;
constexpr uint32_t OFF_F32 = OFF_HALF + 2 * H_BYTES;               // fp32 biases etc. (tc::F_COUNT floats)
constexpr uint32_t OFF_MISC = OFF_F32 + tc::F_RESERVE * 4;         // mbarriers, TMEM slot
constexpr uint32_t SMEM_BYTES = OFF_MISC + 128;
static_assert(H_BYTES % 16 == 0 && OFF_F32 % 16 == 0, "alignment");
static_assert(SMEM_BYTES <= 227 * 1024, "pp64 update kernel exceeds shared memory");

// ---- TMEM columns (fp32); every accumulator is M = 64 (lanes 0-15 of each quadrant)
constexpr uint32_t C_HALF = 80;    // per half: 64 layer-output columns + 16 d(features) columns
constexpr uint32_t C_DF_ = 64;
constexpr uint32_t C_G0 = 160;     // enc0  dW^T [64 out][8 in | bias | 0 x 7]
constexpr uint32_t C_G1 = 176;     // enc1  dW^T [64][64 | bias, 0 x 7]
constexpr uint32_t C_G2 = 248;     // enc2  dW   [64 in][16 out]
constexpr uint32_t C_GV0 = 264;    // v0    dW^T [64][16 | bias, 0 x 7]
constexpr uint32_t C_GV1 = 288;    // v1    dW^T [64][64 | bias, 0 x 7]
constexpr uint32_t C_GA0 = 360;    // a0    dW^T [64][16 | bias, 0 x 7]
constexpr uint32_t C_GA1 = 384;    // a1    dW   [64 in][16 out]
constexpr uint32_t TMEM_COLS = 512;
constexpr int NSTEPS = 12;         // issue steps per half tile

struct Pp64Args {
  const float* params;
  LrxBatchView buf;
  const int32_t* idx;
  long long B;
  double B_global;
  const double* adv_sums;
  LrxPpoHyper h;
  float* partial;           // [gridDim.x][partial_stride]
  long long partial_stride;
  double* loss_partial;     // [gridDim.x][8]
  int n_tiles;              // 64-row tiles
  int obs_dim, n_out, is_box;
  tc::LayerOffsets lo;
  long long* prof;  // test hook (lrx_debug_set_pp64_prof): raw clock64 stamps of CTA 0 / half 0, 8 per step
  int prof_records;
};

__device__ __forceinline__ void tie_weights_f(float a, float b, float& wa, float& wb) {
  wa = a < b ? 1.f : (a == b ? 0.5f : 0.f);  // JAX: d min(a,b) splits ties 1/2-1/2
  wb = 1.f - wa;
}
__device__ __forceinline__ float clip_grad_f(float x, float lo, float hi) {
  return (x > lo && x < hi) ? 1.f : ((x == lo || x == hi) ? 0.5f : 0.f);
}

// tcgen05.ld 16x256b: 16 TMEM lanes; thread t holds lanes r0 = t / 4 and r1 = r0 + 8, columns c = 2 (t % 4) of every
// 8-column chunk j: v[4j], v[4j+1] = (r0, 8j + c), (r0, 8j + c + 1);  v[4j+2], v[4j+3] = (r1, 8j + c), (r1, 8j + c + 1).
__device__ __forceinline__ void tmem_ld_frag32(uint32_t taddr, float* v) {  // 32 columns
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_frag8(uint32_t taddr, float* v) {  // 8 columns
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(done)
      : "r"(umma::smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void half_barrier(int h) {  // the 256 threads of one half (named barriers 1 and 2)
  asm volatile("bar.sync %0, %1;" ::"r"(h + 1), "n"(HTHREADS) : "memory");
}

// NOUT: number of action-head outputs when known at compile time (1 = scalar Box, 2 = CartPole, 3 = Acrobot /
// MountainCar); 0 = read it from the arguments (up to LRX_MAX_NOUT).
template <int NOUT>
__global__ void __launch_bounds__(THREADS, 1) ppo_grad_pp64_kernel(const __grid_constant__ Pp64Args a) {
  constexpr int NO = NOUT ? NOUT : LRX_MAX_NOUT;
  const int n_out = NOUT ? NOUT : a.n_out;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = umma::warp_idx_sync(), lane = tid & 31;
  const bool worker = warp < WORKERS / 32;  // warp-uniform
  const int h = worker ? (warp >> 3) : 0;   // half
  const int wh = warp & 7, q = wh & 3, ch = wh >> 2;
  const int p = lane & 3, r8 = lane >> 2;
  const int ra = 16 * q + r8, rb = ra + 8;  // this thread's two rows of a 64-wide layer's fragment (rows of the half tile)
  const int cb = 32 * ch + 2 * p;           // its first column; chunk j adds 8 j
  float* fp = reinterpret_cast<float*>(smem + OFF_F32);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_MISC);  // req[2], barA[2], barB[2]
  uint64_t* req = bars + h;
  uint64_t* barA = bars + 2 + h;   // critical GEMM(s) of a step done
  uint64_t* barB = bars + 4 + h;   // background GEMMs of a step done
  uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + OFF_MISC + 64);

  uint8_t* hb = smem + OFF_HALF + (uint32_t)h * H_BYTES;
  float* vpart = reinterpret_cast<float*>(hb + H_VPART);
  int* idx_s = reinterpret_cast<int*>(hb + H_IDX);
  float* obs_s = reinterpret_cast<float*>(hb + H_OBS32);
  float* rows_s = reinterpret_cast<float*>(hb + H_PLANES);  // current (read by this tile) / next (being gathered)
  float* rows_n = rows_s + PLANES_BYTES / 4;
  uint8_t* valid_s = hb + H_VALID;
  uint8_t* valid_n = valid_s + HROWS;

  const tc2::Mat W1{smem + OFF_W1, 64, 8192}, W2{smem + OFF_W2, 16, 2048};
  const tc2::Mat WV0{smem + OFF_WV0, 64, 2048}, WV1{smem + OFF_WV1, 64, 8192};
  const tc2::Mat WA0{smem + OFF_WA0, 64, 2048}, WA1{smem + OFF_WA1, 16, 2048};
  const tc2::Mat XOBS{hb + H_XOBS, HROWS, 2 * CG};
  const tc2::Mat XH1{hb + H_XH1, HROWS, 9 * CG}, XH2{hb + H_XH2, HROWS, 8 * CG};
  const tc2::Mat DH1 = XH2;  // d(loss)/d(h1) of a tile reuses the XH2 buffer
  const tc2::Mat XF{hb + H_XF, HROWS, 3 * CG}, XB1{hb + H_XB1, HROWS, 9 * CG};
  const tc2::Mat DY64{hb + H_DY64, HROWS, 8 * CG}, DY16{hb + H_DY16, HROWS, 2 * CG};

  // ------------------------------------------------------------------ one-time staging (as in lrx_update_tc128.cu)
  {
    constexpr int NCH = 512 + 128 + 128 + 512 + 128 + 128, PER = (NCH + THREADS - 1) / THREADS;
    const uint32_t base[6] = {OFF_W1, OFF_W2, OFF_WV0, OFF_WV1, OFF_WA0, OFF_WA1}, lo_off[6] = {8192, 2048, 2048, 8192, 2048, 2048};
    const int lay[6] = {tc::L_ENC1, tc::L_ENC2, tc::L_V0, tc::L_V1, tc::L_A0, tc::L_A1};
    const int outs[6] = {64, 16, 64, 64, 64, 16}, ins[6] = {64, 64, 16, 64, 16, 64};
    const int first[7] = {0, 512, 640, 768, 1280, 1408, 1536};
    float v[PER][8];
    int mi[PER], oo[PER], cc[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int e = tid + k * THREADS;
      mi[k] = -1;
      if (e < NCH) {
        int m6 = 0;
#pragma unroll
        for (int j = 1; j < 6; ++j) m6 += (e >= first[j]) ? 1 : 0;
        const int el = e - first[m6];
        mi[k] = m6; oo[k] = el % outs[m6]; cc[k] = el / outs[m6];
        const float* src = a.params + a.lo.w[lay[m6]] + (size_t)oo[k] * ins[m6] + cc[k] * 8;
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
          const float4 x = __ldg(reinterpret_cast<const float4*>(src)), y = __ldg(reinterpret_cast<const float4*>(src) + 1);
          v[k][0] = x.x; v[k][1] = x.y; v[k][2] = x.z; v[k][3] = x.w; v[k][4] = y.x; v[k][5] = y.y; v[k][6] = y.z; v[k][7] = y.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[k][j] = __ldg(src + j);
        }
      }
    }
    tc::stage_f32(a.params, a.lo, n_out, a.obs_dim, fp, tid, THREADS);
#pragma unroll
    for (int k = 0; k < PER; ++k)
      if (mi[k] >= 0) tc2::store8(tc2::Mat{smem + base[mi[k]], (uint32_t)outs[mi[k]], lo_off[mi[k]]}, cc[k], oo[k], v[k]);
  }
  if (tid < 2 * HROWS) {  // the constant ones columns of both halves (bf16 1.0 = 0x3F80 in element 0 of the extra c-group)
    const uint4 one = make_uint4(0x00003F80u, 0u, 0u, 0u);
    uint8_t* b2 = smem + OFF_HALF + (uint32_t)(tid >> 6) * H_BYTES;
    const uint32_t r = (uint32_t)(tid & 63) * 16u;
    *reinterpret_cast<uint4*>(b2 + H_XOBS + 1 * CG + r) = one;
    *reinterpret_cast<uint4*>(b2 + H_XH1 + 8 * CG + r) = one;
    *reinterpret_cast<uint4*>(b2 + H_XF + 2 * CG + r) = one;
    *reinterpret_cast<uint4*>(b2 + H_XB1 + 8 * CG + r) = one;
  }
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) umma::mbar_init(bars + i, 1);
    umma::fence_mbar_init();
  }
  if (warp == 0) umma::tmem_alloc(tslot, TMEM_COLS);
  umma::fence_async_smem();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tb = *tslot;
  const uint32_t tw = umma::tmem_addr(tb, q * 32, 0);        // this warp's quadrant
  const uint32_t td0 = tw + (uint32_t)h * C_HALF;            // this half's layer-output accumulator
  const uint32_t tdf = td0 + C_DF_;                          // ... and its d(features) accumulator
  if (worker) {  // zero all of TMEM: warp (quadrant warp & 3, column quarter warp >> 2)
    float z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0.f;
    const uint32_t tq = umma::tmem_addr(tb, (warp & 3) * 32, 0);
    for (int c = (warp >> 2) * 128; c < (warp >> 2) * 128 + 128; c += 16) umma::tmem_st16(tq + c, z);
    umma::tmem_st_wait();
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();

  // this half's 64-row tiles: first + k * stride
  const int t_first = 2 * (int)blockIdx.x + h, t_stride = 2 * (int)gridDim.x;

  if (!worker) {
    // ================================================================ issuer warp: serves both halves' requests
    int left[2], st[2] = {0, 0}, irec = 0;
    uint32_t phR[2] = {0u, 0u};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int f = 2 * (int)blockIdx.x + k;
      left[k] = f < a.n_tiles ? (a.n_tiles - f + t_stride - 1) / t_stride : 0;
    }
    while (left[0] > 0 || left[1] > 0) {
      bool served = false;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (left[k] == 0) continue;
        if (!__all_sync(0xffffffffu, mbar_test(bars + k, phR[k]))) continue;
        phR[k] ^= 1u;
        served = true;
        const bool iprof = a.prof != nullptr && blockIdx.x == 0 && k == 0 && lane == 0 && irec < a.prof_records;
        if (iprof) a.prof[irec * 8 + 2] = clock64();
        umma::tc_fence_after();
        const bool ld = umma::elect_one();
        uint8_t* b2 = smem + OFF_HALF + (uint32_t)k * H_BYTES;
        const tc2::Mat xobs{b2 + H_XOBS, HROWS, 2 * CG}, xh1{b2 + H_XH1, HROWS, 9 * CG}, xh2{b2 + H_XH2, HROWS, 8 * CG};
        const tc2::Mat xf{b2 + H_XF, HROWS, 3 * CG}, xb1{b2 + H_XB1, HROWS, 9 * CG};
        const tc2::Mat dy64{b2 + H_DY64, HROWS, 8 * CG}, dy16{b2 + H_DY16, HROWS, 2 * CG};
        const uint32_t d0 = tb + (uint32_t)k * C_HALF, df = d0 + C_DF_;
        // weight-gradient accumulators (M = 64: lanes 0-15 of each quadrant): half 1 accumulates in lanes 16-31 of the same
        // columns, so the two halves never add into one accumulator in arrival order (bitwise reproducible); the flush
        // adds the two in a fixed order
        const uint32_t tg = tb + ((uint32_t)k << 20);
        uint64_t* bA = bars + 2 + k;
        uint64_t* bB = bars + 4 + k;
        // K of a weight-gradient GEMM = the 64 rows of the half tile = 4 slices of 16
        switch (st[k]) {
          case 0:  // encoder layer 1
            tc2::gemm<false, false>(ld, d0, xh1, W1, HROWS, 64, 64, 4, 0u);
            if (ld) umma::mma_commit(bA);
            break;
          case 1:  // encoder output layer
            tc2::gemm<false, false>(ld, d0, xh2, W2, HROWS, 16, 16, 4, 0u);
            if (ld) umma::mma_commit(bA);
            break;
          case 2:  // value head layer 0
            tc2::gemm<false, false>(ld, d0, xf, WV0, HROWS, 64, 64, 1, 0u);
            if (ld) umma::mma_commit(bA);
            break;
          case 3:  // value head layer 1
            tc2::gemm<false, false>(ld, d0, xb1, WV1, HROWS, 64, 64, 4, 0u);
            if (ld) umma::mma_commit(bA);
            break;
          case 4:  // value head layer 1 backward: dX = dY W ; [dW^T | db] += dY^T [X | 1]
            tc2::gemm<false, true>(ld, d0, dy64, WV1, HROWS, 64, 64, 4, 0u);
            if (ld) umma::mma_commit(bA);
            tc2::gemm<true, true>(ld, tg + C_GV1, dy64, xb1, 64, 72, 64, 4, 1u);
            if (ld) umma::mma_commit(bB);
            break;
          case 5:  // action head layer 0 ; value head layer 0 backward -> d(features), its weight gradient
            tc2::gemm<false, false>(ld, d0, xf, WA0, HROWS, 64, 64, 1, 0u);
            if (ld) umma::mma_commit(bA);
            tc2::gemm<false, true>(ld, df, dy64, WV0, HROWS, 16, 16, 4, 0u);
            tc2::gemm<true, true>(ld, tg + C_GV0, dy64, xf, 64, 24, 16, 4, 1u);
            break;
          case 6:  // action head layer 1
            tc2::gemm<false, false>(ld, d0, xb1, WA1, HROWS, 16, 16, 4, 0u);
            if (ld) umma::mma_commit(bA);
            break;
          case 7:  // action head layer 1 backward
            tc2::gemm<false, true>(ld, d0, dy16, WA1, HROWS, 64, 64, 1, 0u);
            if (ld) umma::mma_commit(bA);
            tc2::gemm<true, true>(ld, tg + C_GA1, xb1, dy16, 64, 16, 16, 4, 1u);
            break;
          case 8:  // action head layer 0 backward: d(features) += ...
            tc2::gemm<false, true>(ld, df, dy64, WA0, HROWS, 16, 16, 4, 1u);
            if (ld) umma::mma_commit(bA);
            tc2::gemm<true, true>(ld, tg + C_GA0, dy64, xf, 64, 24, 16, 4, 1u);
            break;
          case 9:  // encoder output layer backward
            tc2::gemm<false, true>(ld, d0, dy16, W2, HROWS, 64, 64, 1, 0u);
            if (ld) umma::mma_commit(bA);
            tc2::gemm<true, true>(ld, tg + C_G2, xh2, dy16, 64, 16, 16, 4, 1u);
            break;
          case 10:  // encoder layer 1 backward
            tc2::gemm<false, true>(ld, d0, dy64, W1, HROWS, 64, 64, 4, 0u);
            if (ld) umma::mma_commit(bA);
            tc2::gemm<true, true>(ld, tg + C_G1, dy64, xh1, 64, 72, 64, 4, 1u);
            break;
          default:  // 11: encoder layer 0 weight gradient [dW^T | db]: obs c-group + the ones c-group; its commit covers
                    // every earlier GEMM of this half (the tensor pipe executes in issue order)
            tc2::gemm<true, true>(ld, tg + C_G0, xh2, xobs, 64, 16, 8, 4, 1u);
            if (ld) umma::mma_commit(bB);
            break;
        }
        if (iprof) a.prof[irec * 8 + 3] = clock64();
        if (k == 0) ++irec;
        __syncwarp();
        if (++st[k] == NSTEPS) {
          st[k] = 0;
          --left[k];
        }
      }
      if (!served) __nanosleep(20);
    }
  } else {
    // ================================================================ epilogue warps of half h
    uint32_t phA = 0, phB = 0;
    bool pendA = false, pendB = false;
    auto wait_crit = [&]() {
      if (pendA) {
        umma::mbar_wait(barA, phA);
        phA ^= 1u;
        pendA = false;
        umma::tc_fence_after();
      }
    };
    auto wait_bg = [&]() {
      if (pendB) {
        umma::mbar_wait(barB, phB);
        phB ^= 1u;
        pendB = false;
        umma::tc_fence_after();
      }
    };
    // publish this half's shared-memory operands and ask the issuer for the next step's GEMMs
    int wrec = -1;
    const bool wprof_thread = a.prof != nullptr && blockIdx.x == 0 && h == 0 && wh == 0 && lane == 0;
    auto mark = [&](int f) {
      if (wprof_thread && wrec >= 0 && wrec < a.prof_records) a.prof[wrec * 8 + f] = clock64();
    };
    auto request = [&](bool crit, bool track_bg) {
      mark(7);
      ++wrec;
      mark(0);
      wait_crit();
      wait_bg();
      umma::fence_async_smem();
      umma::tc_fence_before();
      half_barrier(h);
      if (wh == 0 && lane == 0) umma::mbar_arrive(req);
      mark(1);
      pendA = crit;
      pendB = track_bg;
    };
    const LrxPpoHyper hy = a.h;
    const float invB = (float)(1.0 / a.B_global);
    float adv_mean = 0.f, adv_den = 1.f;
    if (hy.normalize_advantages) {
      const double mean = a.adv_sums[0] / a.B_global;
      double var = a.adv_sums[1] / a.B_global - mean * mean;
      var = var > 0 ? var : 0;
      adv_mean = (float)mean;
      adv_den = (float)sqrt(var) + 1.1920928955078125e-07f;  // + finfo(float32).eps   (ppo.py:424-427)
    }
    float log_std = 0.f, scale = 1.f, log_scale = 0.f;
    if (a.is_box) {
      log_std = __ldg(a.params + a.lo.log_std);
      scale = expf(log_std);
      log_scale = logf(scale);
    }
    float s_kl = 0, s_pol = 0, s_val = 0, s_ent = 0, s_dls = 0, s_bad = 0;

    // ---- epilogue building blocks: rows ra / rb, columns cb + 8 j + {0, 1} of a 64-wide layer
    auto store_frag = [&](const tc2::Mat& dst, const float* v) {
      uint8_t* d = dst.ptr + (uint32_t)(4 * ch) * CG + (uint32_t)ra * 16u + (uint32_t)p * 4u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t ha, la, hb2, lb2;
        tc2::split2(v[4 * j], v[4 * j + 1], ha, la);
        tc2::split2(v[4 * j + 2], v[4 * j + 3], hb2, lb2);
        *reinterpret_cast<uint32_t*>(d + j * CG) = ha;
        *reinterpret_cast<uint32_t*>(d + j * CG + 128) = hb2;  // row rb = ra + 8
        *reinterpret_cast<uint32_t*>(d + j * CG + dst.lo_off) = la;
        *reinterpret_cast<uint32_t*>(d + j * CG + dst.lo_off + 128) = lb2;
      }
    };
    auto add_bias_relu = [&](float* v, const float* bias, uint32_t& mask) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 b = *reinterpret_cast<const float2*>(bias + cb + 8 * j);
        v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.x; v[4 * j + 3] += b.y;
      }
      mask = 0;  // element e at bit 15 - e, 1 = negative pre-activation = no gradient
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        mask = __funnelshift_l(__float_as_uint(v[e]), mask, 1);
        v[e] = fmaxf(v[e], 0.f);
      }
    };
    auto fwd64 = [&](const float* bias, const tc2::Mat& dst, uint32_t& mask) {
      wait_crit();
      mark(4);
      float v[16];
      tmem_ld_frag32(td0 + 32 * ch, v);
      umma::tmem_ld_wait();
      mark(5);
      add_bias_relu(v, bias, mask);
      mark(6);
      store_frag(dst, v);
    };
    auto bwd64 = [&](uint32_t mask, const tc2::Mat& dst, bool need_bg) {
      wait_crit();
      mark(4);
      float v[16];
      tmem_ld_frag32(td0 + 32 * ch, v);
      umma::tmem_ld_wait();
      mark(5);
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e] = ((mask >> (15 - e)) & 1u) ? 0.f : v[e];
      if (need_bg) wait_bg();
      mark(6);
      store_frag(dst, v);
    };
    // 16-wide layer outputs (features, d(features)): warp (q, ch) takes columns 8 ch .. 8 ch + 7 of its quadrant's rows
    float bs_e2[2] = {0.f, 0.f};  // encoder output layer's bias gradient: columns 8 ch + 2 p + {0, 1} over this thread's rows
    auto half16 = [&](uint32_t taddr, const float* bias, const tc2::Mat& dst, bool sum_cols) {
      wait_crit();
      float v[4];
      tmem_ld_frag8(taddr + 8 * ch, v);
      umma::tmem_ld_wait();
      if (bias) {
        const float2 b = *reinterpret_cast<const float2*>(bias + 8 * ch + 2 * p);
        v[0] += b.x; v[1] += b.y; v[2] += b.x; v[3] += b.y;
      }
      if (sum_cols) {
        bs_e2[0] += v[0] + v[2];
        bs_e2[1] += v[1] + v[3];
      }
      uint32_t ha, la, hb2, lb2;
      tc2::split2(v[0], v[1], ha, la);
      tc2::split2(v[2], v[3], hb2, lb2);
      uint8_t* d = dst.ptr + (uint32_t)ch * CG + (uint32_t)ra * 16u + (uint32_t)p * 4u;
      *reinterpret_cast<uint32_t*>(d) = ha;
      *reinterpret_cast<uint32_t*>(d + 128) = hb2;
      *reinterpret_cast<uint32_t*>(d + dst.lo_off) = la;
      *reinterpret_cast<uint32_t*>(d + dst.lo_off + 128) = lb2;
    };

    // ---- row gather (buffer/base_buffer.py:90-103): cp.async straight into shared staging, one tile ahead
    auto cp_async4 = [](void* dst, const void* src, bool ok) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(umma::smem_u32(dst)), "l"(src), "r"(ok ? 4 : 0) : "memory");
    };
    auto cp_async16 = [](void* dst, const void* src, bool ok) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(umma::smem_u32(dst)), "l"(src), "r"(ok ? 16 : 0) : "memory");
    };
    auto tile_row_ok = [&](int t, int r) { return t < a.n_tiles && (long long)t * HROWS + r < a.B; };
    auto gather_row = [&](int r, int i, bool ok) {  // buffer row i (zeros when !ok) -> staging slot r
      if (a.buf.packed != nullptr) {
        const float* rec = a.buf.packed + (ok ? (size_t)i * LRX_PACKED_FLOATS : 0);
        cp_async16(obs_s + r * 8, rec, ok);
        cp_async16(obs_s + r * 8 + 4, rec + 4, ok);
        cp_async4(rows_n + R_ACT * HROWS + r, rec + 8, ok);
        cp_async4(rows_n + R_LOGP * HROWS + r, rec + 9, ok);
        cp_async4(rows_n + R_VAL * HROWS + r, rec + 10, ok && hy.clip_value_loss);
        cp_async4(rows_n + R_RET * HROWS + r, rec + 11, ok);
        cp_async4(rows_n + R_ADV * HROWS + r, rec + 12, ok);
        valid_n[r] = ok ? 1 : 0;
        return;
      }
      const int n = i / a.buf.T, t = i - n * a.buf.T;
      const size_t src = ok ? (size_t)t * a.buf.N + n : 0;
      cp_async4(rows_n + R_RET * HROWS + r, a.buf.ret + src, ok);
      cp_async4(rows_n + R_VAL * HROWS + r, (hy.clip_value_loss ? a.buf.val : a.buf.ret) + src, ok && hy.clip_value_loss);
      cp_async4(rows_n + R_ACT * HROWS + r, (const int32_t*)a.buf.act + src, ok);
      cp_async4(rows_n + R_LOGP * HROWS + r, a.buf.logp + src, ok);
      cp_async4(rows_n + R_ADV * HROWS + r, a.buf.adv + src, ok);
      const float* op = a.buf.obs + src * a.obs_dim;
      if (a.obs_dim == 4) {
        cp_async16(obs_s + r * 8, op, ok);
        *reinterpret_cast<float4*>(obs_s + r * 8 + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) cp_async4(obs_s + r * 8 + k, op + (k < a.obs_dim ? k : 0), ok && k < a.obs_dim);
      }
      valid_n[r] = ok ? 1 : 0;
    };
    const bool gatherer = wh >= 4 && wh < 6;   // two warps = 64 lanes = the 64 rows of the next tile
    const int gr = (wh - 4) * 32 + lane;
    auto gather_tile = [&](int t_next, int t_after, bool first) {
      const bool ok = tile_row_ok(t_next, gr);
      const int i = first ? (ok ? __ldg(a.idx + (long long)t_next * HROWS + gr) : 0) : idx_s[gr];
      gather_row(gr, i, ok);
      const bool ok2 = tile_row_ok(t_after, gr);
      cp_async4(idx_s + gr, a.idx + (ok2 ? (long long)t_after * HROWS + gr : 0), ok2);
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto gather_wait = []() { asm volatile("cp.async.wait_group 0;" ::: "memory"); };
    auto swap_staging = [&]() {
      float* t = rows_s; rows_s = rows_n; rows_n = t;
      uint8_t* u = valid_s; valid_s = valid_n; valid_n = u;
    };
    if (gatherer) {
      gather_tile(t_first, t_first + t_stride, true);
      gather_wait();
    }
    swap_staging();
    half_barrier(h);

    // gradients kept in registers across the tiles of the half
    float gw_v2[8], gb_v2 = 0.f;  // value output layer (64 -> 1): dW of columns cb + 8 j + {0, 1} over this thread's rows; db
#pragma unroll
    for (int i = 0; i < 8; ++i) gw_v2[i] = 0.f;
    // action_dist.mapping (16 -> n_out): dW[o][f_c] over this lane's row of every tile, db[o]; f = {2p, 2p+1, 2p+8, 2p+9}
    float gm_w[NO][4], gm_b[NO];
#pragma unroll
    for (int o2 = 0; o2 < NO; ++o2) gm_w[o2][0] = gm_w[o2][1] = gm_w[o2][2] = gm_w[o2][3] = gm_b[o2] = 0.f;
    const bool owner = ch == 0 && p == 0;  // one thread per row pair (ra, rb): per-row statistics

    for (int tile = t_first; tile < a.n_tiles; tile += t_stride) {
      const bool valid_a = valid_s[ra] != 0, valid_b = valid_s[rb] != 0;

      // -------------------------------------------------------------- encoder layer 0 (d <= 8 inputs) on the FMA pipe
      uint32_t mk_h1 = 0, mk_h2 = 0, mk_b1 = 0, mk_b2 = 0;
      {
        float xa[8], xb[8];
        {
          const float4 a0 = *reinterpret_cast<const float4*>(obs_s + ra * 8), a1 = *reinterpret_cast<const float4*>(obs_s + ra * 8 + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(obs_s + rb * 8), b1 = *reinterpret_cast<const float4*>(obs_s + rb * 8 + 4);
          xa[0] = a0.x; xa[1] = a0.y; xa[2] = a0.z; xa[3] = a0.w; xa[4] = a1.x; xa[5] = a1.y; xa[6] = a1.z; xa[7] = a1.w;
          xb[0] = b0.x; xb[1] = b0.y; xb[2] = b0.z; xb[3] = b0.w; xb[4] = b1.x; xb[5] = b1.y; xb[6] = b1.z; xb[7] = b1.w;
        }
        float v[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) {
            const float* wp = fp + tc::F_W0 + (cb + 8 * j + c2) * 8;
            const float4 wa = *reinterpret_cast<const float4*>(wp);
            float sa = fmaf(wa.w, xa[3], fmaf(wa.z, xa[2], fmaf(wa.y, xa[1], wa.x * xa[0])));
            float sb = fmaf(wa.w, xb[3], fmaf(wa.z, xb[2], fmaf(wa.y, xb[1], wa.x * xb[0])));
            if (a.obs_dim > 4) {
              const float4 wb = *reinterpret_cast<const float4*>(wp + 4);
              sa = fmaf(wb.w, xa[7], fmaf(wb.z, xa[6], fmaf(wb.y, xa[5], fmaf(wb.x, xa[4], sa))));
              sb = fmaf(wb.w, xb[7], fmaf(wb.z, xb[6], fmaf(wb.y, xb[5], fmaf(wb.x, xb[4], sb))));
            }
            v[4 * j + c2] = sa;
            v[4 * j + 2 + c2] = sb;
          }
        }
        add_bias_relu(v, fp + tc::F_B0, mk_h1);
        wait_crit();
        wait_bg();  // the previous tile's last GEMMs read XH1 / XH2 / XOBS
        store_frag(XH1, v);
        if (owner) {
          float z8[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) z8[k] = k < a.obs_dim ? xa[k] : 0.f;
          tc2::store8(XOBS, 0, ra, z8);
#pragma unroll
          for (int k = 0; k < 8; ++k) z8[k] = k < a.obs_dim ? xb[k] : 0.f;
          tc2::store8(XOBS, 0, rb, z8);
        }
      }
      request(true, false);                                   // step 0: enc1
      fwd64(fp + tc::F_B1, XH2, mk_h2);
      request(true, false);                                   // step 1: enc2
      half16(td0, fp + tc::F_B2, XF, false);
      request(true, false);                                   // step 2: v0
      fwd64(fp + tc::F_BV0, XB1, mk_b1);
      request(true, false);                                   // step 3: v1
      // -------------------------------------------------------------- value head output (64 -> 1) on the FMA pipe
      float xb2[16];
      {
        wait_crit();
        tmem_ld_frag32(td0 + 32 * ch, xb2);
        umma::tmem_ld_wait();
        add_bias_relu(xb2, fp + tc::F_BV1, mk_b2);
        float pa = 0.f, pb = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 w = *reinterpret_cast<const float2*>(fp + tc::F_WV2 + cb + 8 * j);
          pa = fmaf(w.y, xb2[4 * j + 1], fmaf(w.x, xb2[4 * j], pa));
          pb = fmaf(w.y, xb2[4 * j + 3], fmaf(w.x, xb2[4 * j + 2], pb));
        }
        pa += __shfl_xor_sync(0xffffffffu, pa, 1); pa += __shfl_xor_sync(0xffffffffu, pa, 2);
        pb += __shfl_xor_sync(0xffffffffu, pb, 1); pb += __shfl_xor_sync(0xffffffffu, pb, 2);
        if (p == 0) { vpart[ch * HROWS + ra] = pa; vpart[ch * HROWS + rb] = pb; }
      }
      half_barrier(h);
      // -------------------------------------------------------------- value loss (ppo.py:435-451) and its backward
      {
        auto vloss = [&](int r, bool valid, float& dv, float& sv, float& value) {
          value = (fp[tc::F_BV2] + vpart[r]) + vpart[HROWS + r];
          dv = 0.f; sv = 0.f;
          if (valid) {
            const float ret_r = rows_s[R_RET * HROWS + r], vold_r = rows_s[R_VAL * HROWS + r];
            const float c = hy.clip_coefficient;
            const float e1 = value - ret_r;
            float dvl;
            if (hy.clip_value_loss) {
              const float dvo = value - vold_r;
              const float clipped = vold_r + lrx_clipf(dvo, -c, c);
              const float e2 = clipped - ret_r;
              const float q1 = e1 * e1, q2 = e2 * e2;
              sv = fminf(q1, q2);
              float u1, u2;
              tie_weights_f(q1, q2, u1, u2);
              dvl = invB * (u1 * e1 + u2 * e2 * clip_grad_f(dvo, -c, c));
            } else {
              sv = e1 * e1;
              dvl = invB * e1;
            }
            dv = hy.value_loss_coefficient * dvl;
          }
        };
        float dva, dvb, sva, svb, vala, valb;
        vloss(ra, valid_a, dva, sva, vala);
        vloss(rb, valid_b, dvb, svb, valb);
        if (owner) {
          if (valid_a) { if (!isfinite(vala)) s_bad += 1.0f; s_val += sva; gb_v2 += dva; }
          if (valid_b) { if (!isfinite(valb)) s_bad += 1.0f; s_val += svb; gb_v2 += dvb; }
        }
        float v[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 w = *reinterpret_cast<const float2*>(fp + tc::F_WV2 + cb + 8 * j);
          gw_v2[2 * j] = fmaf(xb2[4 * j + 2], dvb, fmaf(xb2[4 * j], dva, gw_v2[2 * j]));
          gw_v2[2 * j + 1] = fmaf(xb2[4 * j + 3], dvb, fmaf(xb2[4 * j + 1], dva, gw_v2[2 * j + 1]));
          v[4 * j] = dva * w.x; v[4 * j + 1] = dva * w.y; v[4 * j + 2] = dvb * w.x; v[4 * j + 3] = dvb * w.y;
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = ((mk_b2 >> (15 - e)) & 1u) ? 0.f : v[e];
        store_frag(DY64, v);
      }
      request(true, true);                                    // step 4: v1 bwd (+ dW v1, tracked)
      bwd64(mk_b1, DY64, true);                               // the background GEMM reads DY64 and XB1
      request(true, false);                                   // step 5: a0 (+ dF from the value head, dW v0)
      fwd64(fp + tc::F_BA0, XB1, mk_b1);
      request(true, false);                                   // step 6: a1
      // the next tile's rows, fetched under the policy loss
      if (gatherer) gather_tile(tile + t_stride, tile + 2 * t_stride, false);
      {
        // ------------------------------------------------------------ action head output, policy loss and their backward:
        // four lanes per row.  Warp (q, ch) serves rows 16 q + 8 ch + lane / 4; lane p holds features 2p, 2p+1, 2p+8, 2p+9.
        const int rr = 16 * q + 8 * ch + r8;
        float x[4];
        {
          float a2[8];
          wait_crit();
          umma::tmem_ld_frag16(td0, a2);
          umma::tmem_ld_wait();
          x[0] = ch ? a2[2] : a2[0]; x[1] = ch ? a2[3] : a2[1]; x[2] = ch ? a2[6] : a2[4]; x[3] = ch ? a2[7] : a2[5];
          const float2 b0 = *reinterpret_cast<const float2*>(fp + tc::F_BA1 + 2 * p);
          const float2 b1 = *reinterpret_cast<const float2*>(fp + tc::F_BA1 + 2 * p + 8);
          x[0] += b0.x; x[1] += b0.y; x[2] += b1.x; x[3] += b1.y;
        }
        float head[NO], wm[NO][4];
#pragma unroll
        for (int o2 = 0; o2 < NO; ++o2) {
          float sacc = 0.f;
          if (o2 < n_out) {
            const float2 w0 = *reinterpret_cast<const float2*>(fp + tc::F_WMAP + o2 * 16 + 2 * p);
            const float2 w1 = *reinterpret_cast<const float2*>(fp + tc::F_WMAP + o2 * 16 + 2 * p + 8);
            wm[o2][0] = w0.x; wm[o2][1] = w0.y; wm[o2][2] = w1.x; wm[o2][3] = w1.y;
            sacc = fmaf(w1.y, x[3], fmaf(w1.x, x[2], fmaf(w0.y, x[1], w0.x * x[0])));
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
            sacc += fp[tc::F_BMAP + o2];
          } else {
            wm[o2][0] = wm[o2][1] = wm[o2][2] = wm[o2][3] = 0.f;
          }
          head[o2] = sacc;
        }
        float dhead[NO];
#pragma unroll
        for (int o2 = 0; o2 < NO; ++o2) dhead[o2] = 0.f;
        if (valid_s[rr] != 0) {
          const int act_raw = reinterpret_cast<const int*>(rows_s)[R_ACT * HROWS + rr];
          const float logp_old = rows_s[R_LOGP * HROWS + rr], A = rows_s[R_ADV * HROWS + rr];
          float logp, entropy, z_n = 0.f;
          float lp[NO], pr[NO];
          if (a.is_box) {
            const float loc = head[0];
            z_n = (__int_as_float(act_raw) - loc) / scale;
            logp = -0.5f * (z_n * z_n) - (LRX_HALF_LOG_2PI + log_scale);
            entropy = 0.5f + LRX_HALF_LOG_2PI + log_scale;
          } else {
#pragma unroll
            for (int j = 0; j < NO; ++j) lp[j] = j < n_out ? head[j] : 0.f;
            log_softmax_n<NO>(lp, n_out);
            logp = 0.f;
            entropy = 0.f;
#pragma unroll
            for (int j = 0; j < NO; ++j) {
              if (j < n_out) {
                pr[j] = expf(lp[j]);
                entropy -= pr[j] * lp[j];
                if (j == act_raw) logp = lp[j];
              }
            }
          }
          const bool bad_p = !(isfinite(logp) && isfinite(entropy));  // ppo.py:413-417
          const float log_ratio = logp - logp_old;
          const float ratio = expf(log_ratio);
          const float c = hy.clip_coefficient, lo = 1.f - c, hi = 1.f + c;
          const float An = hy.normalize_advantages ? (A - adv_mean) / adv_den : A;  // ppo.py:424-427
          float dlogp, spol;
          if (hy.surrogate == 1) {  // A2C / REINFORCE: -mean(logp * A)   (a2c.py:401, reinforce.py:391)
            spol = logp * An;
            dlogp = -invB * An;
          } else {
            const float s1 = An * ratio, s2 = An * lrx_clipf(ratio, lo, hi);
            spol = fminf(s1, s2);
            float w1, w2;
            tie_weights_f(s1, s2, w1, w2);
            dlogp = -invB * (w1 * An * ratio + w2 * An * clip_grad_f(ratio, lo, hi) * ratio);
          }
          const float che = hy.entropy_loss_coefficient;
          float sdls = 0.f;
          if (a.is_box) {
            dhead[0] = dlogp * z_n / scale;
            sdls = dlogp * (z_n * z_n - 1.0f);
          } else {
#pragma unroll
            for (int j = 0; j < NO; ++j) {
              if (j < n_out) {
                const float dH = -pr[j] * (lp[j] + entropy);
                dhead[j] = dlogp * ((j == act_raw ? 1.f : 0.f) - pr[j]) + (-che * invB) * dH;
              }
            }
          }
          if (p == 0) {
            if (bad_p) s_bad += 1.0f;
            s_kl += ratio - log_ratio;
            s_pol += spol;
            s_ent += entropy;
            s_dls += sdls;
#pragma unroll
            for (int o2 = 0; o2 < NO; ++o2) gm_b[o2] += dhead[o2];
          }
        }
        // action head backward (mapping, FMA): this lane's four columns of d(a2), its share of the mapping layer's dW
        float d4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float sacc = 0.f;
#pragma unroll
          for (int o2 = 0; o2 < NO; ++o2) {
            if (o2 < n_out) {
              sacc = fmaf(dhead[o2], wm[o2][c], sacc);
              gm_w[o2][c] = fmaf(dhead[o2], x[c], gm_w[o2][c]);
            }
          }
          d4[c] = sacc;
        }
        uint32_t h0, l0, h1, l1;
        tc2::split2(d4[0], d4[1], h0, l0);   // columns 2p, 2p + 1 (c-group 0)
        tc2::split2(d4[2], d4[3], h1, l1);   // columns 2p + 8, 2p + 9 (c-group 1)
        uint8_t* d = DY16.ptr + (uint32_t)rr * 16u + (uint32_t)p * 4u;
        *reinterpret_cast<uint32_t*>(d) = h0;
        *reinterpret_cast<uint32_t*>(d + CG) = h1;
        *reinterpret_cast<uint32_t*>(d + DY16.lo_off) = l0;
        *reinterpret_cast<uint32_t*>(d + DY16.lo_off + CG) = l1;
      }
      request(true, false);                                   // step 7: a1 bwd (+ dW a1)
      bwd64(mk_b1, DY64, false);
      request(true, false);                                   // step 8: a0 bwd -> d(features) (+ dW a0)
      half16(tdf, nullptr, DY16, true);
      request(true, false);                                   // step 9: enc2 bwd (+ dW enc2)
      bwd64(mk_h2, DY64, false);
      request(true, false);                                   // step 10: enc1 bwd (+ dW enc1)
      bwd64(mk_h1, DH1, false);                               // into the XH2 buffer (dead: read last by step 9's GEMMs)
      if (gatherer) gather_wait();                            // the barrier of the request below publishes the next rows
      request(false, true);                                   // step 11: dW enc0, tracked: covers everything issued so far
      swap_staging();
    }
    wait_crit();
    wait_bg();

    // stash the register-held sums for the flush below (after the CTA barrier; the activation buffers are then free)
    __syncwarp();
    // (kept in registers until the barrier)
    // ---------------------------------------------------------------- handled after the barrier
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();

    float* part = a.partial + (size_t)blockIdx.x * a.partial_stride;
    const tc::LayerOffsets& lo = a.lo;
    // ---- TMEM-held weight gradients: warp (quadrant warp & 3, group warp >> 2) as in lrx_update_tc128.cu
    {
      const int fq = warp & 3, fg = warp >> 2;
      const uint32_t tq = umma::tmem_addr(tb, fq * 32, 0);
      const bool low = lane < 16;
      const int mo = fq * 16 + (lane & 15);
      auto flush_T = [&](uint32_t col, int in_real, int in_pad, int bias_col, int layer) {
        for (int c0 = 0; c0 < in_pad; c0 += 8) {
          float v[8];
          umma::tmem_ld8(tq + col + c0, v);
          umma::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += __shfl_down_sync(0xffffffffu, v[j], 16);  // half 0 + half 1
          if (low) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (c0 + j < in_real) part[lo.w[layer] + mo * in_real + c0 + j] = v[j];
          }
        }
        float v[8];
        umma::tmem_ld8(tq + col + bias_col, v);
        umma::tmem_ld_wait();
        v[0] += __shfl_down_sync(0xffffffffu, v[0], 16);
        if (low) part[lo.b[layer] + mo] = v[0];
      };
      auto flush_N = [&](uint32_t col, int in_real, int out_real, int out_pad, int layer) {
        for (int c0 = 0; c0 < out_pad; c0 += 8) {
          float v[8];
          umma::tmem_ld8(tq + col + c0, v);
          umma::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += __shfl_down_sync(0xffffffffu, v[j], 16);
          if (low) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (c0 + j < out_real && mo < in_real) part[lo.w[layer] + (c0 + j) * in_real + mo] = v[j];
          }
        }
      };
      if (fg == 0) { flush_T(C_G0, a.obs_dim, 8, 8, tc::L_ENC0); flush_N(C_G2, 64, 16, 16, tc::L_ENC2); }
      if (fg == 1) flush_T(C_G1, 64, 64, 64, tc::L_ENC1);
      if (fg == 2) { flush_T(C_GV1, 64, 64, 64, tc::L_V1); flush_T(C_GV0, 16, 16, 16, tc::L_V0); }
      if (fg == 3) { flush_T(C_GA0, 16, 16, 16, tc::L_A0); flush_N(C_GA1, 64, 16, 16, tc::L_A1); }
    }
    // ---- register-held gradients (fixed order).  Scratch = half 0's activation region.
    float* sh = reinterpret_cast<float*>(smem + OFF_HALF);
    float* s_wv2 = sh;                        // [8 slots (half, quadrant)][64 columns]
    float* s_bv2 = s_wv2 + 8 * 64;            // [8 slots]
    float* s_e2 = s_bv2 + 8;                  // [8 slots][16 columns]
    float* s_mw = s_e2 + 8 * 16;              // [16 warps][LRX_MAX_NOUT][16]
    float* s_mb = s_mw + 16 * LRX_MAX_NOUT * 16;  // [16 warps][LRX_MAX_NOUT]
    double* shd = reinterpret_cast<double*>(s_mb + 16 * LRX_MAX_NOUT + 8);  // [6][16 warps]; 8-byte aligned below
    shd = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(shd) + 7) & ~(uintptr_t)7);
    const int slot = h * 4 + q;
    {
      // sums over the eight row pairs of the warp (lanes with equal p): xor 4, 8, 16
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float x = gw_v2[i];
        x += __shfl_xor_sync(0xffffffffu, x, 4);
        x += __shfl_xor_sync(0xffffffffu, x, 8);
        x += __shfl_xor_sync(0xffffffffu, x, 16);
        if (lane < 4) s_wv2[slot * 64 + cb + 8 * (i >> 1) + (i & 1)] = x;
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float x = bs_e2[i];
        x += __shfl_xor_sync(0xffffffffu, x, 4);
        x += __shfl_xor_sync(0xffffffffu, x, 8);
        x += __shfl_xor_sync(0xffffffffu, x, 16);
        if (lane < 4) s_e2[slot * 16 + 8 * ch + 2 * p + i] = x;
      }
      {
        const float x = warp_sum(gb_v2);  // non-owner threads hold 0
        if (ch == 0 && lane == 0) s_bv2[slot] = x;
      }
#pragma unroll
      for (int o2 = 0; o2 < NO; ++o2) {
        if (o2 < n_out) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float x = gm_w[o2][c];
            x += __shfl_xor_sync(0xffffffffu, x, 4);
            x += __shfl_xor_sync(0xffffffffu, x, 8);
            x += __shfl_xor_sync(0xffffffffu, x, 16);
            if (lane < 4) s_mw[(warp * LRX_MAX_NOUT + o2) * 16 + 2 * p + (c & 1) + 8 * (c >> 1)] = x;
          }
          const float xb = warp_sum(gm_b[o2]);
          if (lane == 0) s_mb[warp * LRX_MAX_NOUT + o2] = xb;
        }
      }
      double v6[6] = {(double)s_kl, (double)s_pol, (double)s_val, (double)s_ent, (double)s_dls, (double)s_bad};
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        v6[k] = warp_sum(v6[k]);
        if (lane == 0) shd[k * 16 + warp] = v6[k];
      }
    }
    asm volatile("bar.sync 3, %0;" ::"n"(WORKERS) : "memory");  // the 16 epilogue warps
    auto sum8 = [](const float* ptr, int stride) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += ptr[w * stride];  // slot order: fixed
      return t;
    };
    auto sum16 = [](const float* ptr, int stride) {
      float t = 0.f;
      for (int w = 0; w < 16; ++w) t += ptr[w * stride];  // warp order: fixed
      return t;
    };
    if (tid < 64) part[lo.w[tc::L_V2] + tid] = sum8(s_wv2 + tid, 64);
    if (tid == 64) part[lo.b[tc::L_V2]] = sum8(s_bv2, 1);
    if (tid >= 96 && tid < 112) part[lo.b[tc::L_ENC2] + tid - 96] = sum8(s_e2 + tid - 96, 16);
    if (tid >= 128 && tid < 128 + 16 * LRX_MAX_NOUT) {
      const int o2 = (tid - 128) >> 4, i2 = (tid - 128) & 15;
      if (o2 < n_out) part[lo.w[tc::L_MAP] + o2 * 16 + i2] = sum16(s_mw + o2 * 16 + i2, LRX_MAX_NOUT * 16);
    }
    if (tid >= 256 + 32 && tid < 256 + 32 + LRX_MAX_NOUT && tid - 288 < n_out)
      part[lo.b[tc::L_MAP] + tid - 288] = sum16(s_mb + tid - 288, LRX_MAX_NOUT);
    if (tid >= 320 && tid < 336) {  // db_a1 = W_map^T db_map (the layers are linear in between)
      const int i2 = tid - 320;
      float sacc = 0.f;
      for (int o2 = 0; o2 < n_out; ++o2) sacc = fmaf(sum16(s_mb + o2, LRX_MAX_NOUT), fp[tc::F_WMAP + o2 * 16 + i2], sacc);
      part[lo.b[tc::L_A1] + i2] = sacc;
    }
    if (tid >= 352 && tid < 358) {
      const int k = tid - 352;
      double s = 0;
      for (int w = 0; w < 16; ++w) s += shd[k * 16 + w];
      a.loss_partial[(size_t)blockIdx.x * 8 + k] = s;
    }
  }
  if (!worker) {  // the issuer joins the barrier the epilogue warps passed before the flush
    umma::tc_fence_before();
    __syncthreads();
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tb, TMEM_COLS);
}

struct FusedWorkspace {
  float* partial;        // [LRX_MAX_SMS][P_pad]
  double* loss_partial;  // [LRX_MAX_SMS][8]
  size_t P_pad, total;
};
size_t align_up_(size_t x, size_t al) { return (x + al - 1) / al * al; }
FusedWorkspace carve_fused(const LrxPolicyDesc& pol, void* base) {  // same layout as the other two tensor-core kernels
  FusedWorkspace w{};
  w.P_pad = align_up_((size_t)pol.n_params + LRX_PPO_NSTATS, 32);
  char* b = (char*)base;
  size_t off = 0;
  w.partial = (float*)(b ? b + off : nullptr);
  off += align_up_((size_t)LRX_MAX_SMS * w.P_pad * 4, 256);
  w.loss_partial = (double*)(b ? b + off : nullptr);
  off += align_up_((size_t)LRX_MAX_SMS * 8 * sizeof(double), 256);
  w.total = off;
  return w;
}

}  // namespace

// defined in lrx_update.cu
struct LrxApplyArgs;
int lrx_launch_reduce_partials(const float* partial, int splits, long long stride, int P, const double* loss_partial,
                               int loss_blocks, int log_std_off, float ent_coef, double B_global, float* gradbuf,
                               const LrxApplyArgs* apply, cudaStream_t s);

static long long* g_pp64_prof = nullptr;  // test hook, include/lerax_b200_debug.h
static int g_pp64_prof_records = 0;
extern "C" void lrx_debug_set_pp64_prof(long long* device_buffer, int records) {
  g_pp64_prof = device_buffer;
  g_pp64_prof_records = records;
}

int lrx_ppo_grad_pp64(const LrxPolicyDesc* pol, const float* params, const LrxBatchView* buf, const int32_t* idx,
                      int64_t B, int64_t B_global, const double* adv_sums, const LrxPpoHyper* h, float* gradbuf,
                      void* workspace, size_t workspace_bytes, cudaStream_t stream, bool* handled,
                      const LrxApplyArgs* apply) {
  *handled = false;
  if (!tc::is_default_policy(*pol) || lrx_fused_enabled() != 3) return LRX_OK;
  FusedWorkspace w = carve_fused(*pol, workspace);
  LRX_REQUIRE(workspace_bytes >= w.total, LRX_ERR_INVALID_ARG, "lrx_ppo_grad: workspace too small for the fused path");
  Pp64Args a{};
  a.params = params;
  a.buf = *buf;
  a.idx = idx;
  a.B = B;
  a.B_global = (double)B_global;
  a.adv_sums = adv_sums;
  a.h = *h;
  a.partial = w.partial;
  a.partial_stride = (long long)w.P_pad;
  a.loss_partial = w.loss_partial;
  a.n_tiles = (int)((B + HROWS - 1) / HROWS);
  a.obs_dim = pol->obs_dim;
  a.n_out = pol->n_out;
  a.is_box = pol->is_box;
  a.lo = tc::layer_offsets(*pol);
  a.prof = g_pp64_prof;
  a.prof_records = g_pp64_prof_records;
  const int sms = lrx_sm_count();
  const int pairs = (a.n_tiles + 1) / 2;
  const int grid = pairs < sms ? pairs : sms;
  static bool set0[64], set1[64], set2[64], set3[64];
  LRX_CUDA_CHECK(lrx_set_max_smem(ppo_grad_pp64_kernel<0>, (int)SMEM_BYTES, set0));
  LRX_CUDA_CHECK(lrx_set_max_smem(ppo_grad_pp64_kernel<1>, (int)SMEM_BYTES, set1));
  LRX_CUDA_CHECK(lrx_set_max_smem(ppo_grad_pp64_kernel<2>, (int)SMEM_BYTES, set2));
  LRX_CUDA_CHECK(lrx_set_max_smem(ppo_grad_pp64_kernel<3>, (int)SMEM_BYTES, set3));
  switch (pol->n_out) {
    case 1: ppo_grad_pp64_kernel<1><<<grid, THREADS, SMEM_BYTES, stream>>>(a); break;
    case 2: ppo_grad_pp64_kernel<2><<<grid, THREADS, SMEM_BYTES, stream>>>(a); break;
    case 3: ppo_grad_pp64_kernel<3><<<grid, THREADS, SMEM_BYTES, stream>>>(a); break;
    default: ppo_grad_pp64_kernel<0><<<grid, THREADS, SMEM_BYTES, stream>>>(a); break;
  }
  LRX_LAUNCH_CHECK();
  const int rc = lrx_launch_reduce_partials(w.partial, grid, (long long)w.P_pad, pol->n_params, w.loss_partial, grid,
                                            a.lo.log_std, h->entropy_loss_coefficient, (double)B_global, gradbuf, apply, stream);
  if (rc) return rc;
  LRX_LAUNCH_CHECK();
  *handled = true;
  return LRX_OK;
}
