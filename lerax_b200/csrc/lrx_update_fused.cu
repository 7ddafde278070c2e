// Fused PPO minibatch gradient for the default MLPActorCriticPolicy widths on the tensor
// cores (tcgen05, bf16x3 fp32 emulation, accumulators in TMEM).
//
// Reference: algorithm/ppo.py:396-469 (ppo_loss + filter_value_and_grad) over the gathered
// minibatch of buffer/base_buffer.py:90-103, i.e. exactly what lrx_ppo_grad's generic
// layer-by-layer path computes; this kernel is tried first and handles the default widths
// (encoder d->64->64->16, value head 16->64->64->1, action head 16->64->16->n_out).
//
// One persistent CTA per SM.  The Linear weights (bf16 pieces) stay resident in shared
// memory for the whole launch; a CTA walks 64-row tiles of the minibatch:
//   gather rows -> encoder fwd -> value head fwd -> value loss -> value head bwd ->
//   action head fwd -> policy loss -> action head bwd -> encoder bwd,
// every Linear fwd / dX being one group of tcgen05.mma (M = 64 rows) whose fp32 result is
// pulled from TMEM by the row's thread for bias / ReLU / loss arithmetic and written back as
// the next operand.  The weight and bias gradients (contraction over the rows) are
// tcgen05.mma groups too, ACCUMULATED IN TMEM ACROSS ALL TILES of the CTA and flushed once
// per launch as a per-CTA partial; a fixed-order reduction over CTAs (reduce_partials_kernel)
// makes the result deterministic.  HBM traffic per row is the 32-36 gathered bytes.
#include <stdlib.h>

#include "lrx_common.cuh"
#include "lrx_policy.cuh"
#include "lrx_tc.cuh"

namespace {

constexpr int ROWS = 64;        // rows per tile = UMMA M
constexpr int WORKERS = 512;    // 16 epilogue warps: warp w serves TMEM quadrant q = w & 3 (rows 16q..16q+15 on lanes
                                // 0-15) and column group g = w >> 2 (16 of the 64 columns of a layer output)
constexpr int THREADS = WORKERS + 32;  // + warp 16: issues every tcgen05.mma (converged, one elected lane)

// ---- shared memory map (bytes).  Matrices: piece stride = R * C * 2.
constexpr uint32_t PS16 = ROWS * 16 * 2;  // [64][16] piece
constexpr uint32_t PS64 = ROWS * 64 * 2;  // [64][64] piece
constexpr uint32_t OFF_W0 = 0;                         // second observation buffer [64][16] (encoder layer 0 itself runs on the FMA pipe)
constexpr uint32_t OFF_W1 = OFF_W0 + 3 * PS16;         // [64][64]
constexpr uint32_t OFF_W2 = OFF_W1 + 3 * PS64;         // [16][64]  (piece = 16*64*2 = PS16)
constexpr uint32_t OFF_WV0 = OFF_W2 + 3 * PS16;        // [64][16]
constexpr uint32_t OFF_WV1 = OFF_WV0 + 3 * PS16;       // [64][64]
constexpr uint32_t OFF_WA0 = OFF_WV1 + 3 * PS64;       // [64][16]
constexpr uint32_t OFF_WA1 = OFF_WA0 + 3 * PS16;       // [16][64]
constexpr uint32_t OFF_XOBS = OFF_WA1 + 3 * PS16;      // [64][16] observations (zero padded)
constexpr uint32_t OFF_XH1 = OFF_XOBS + 3 * PS16;      // [64][64]
constexpr uint32_t OFF_XH2 = OFF_XH1 + 3 * PS64;       // [64][64]
constexpr uint32_t OFF_XF = OFF_XH2 + 3 * PS64;        // [64][16] features
constexpr uint32_t OFF_XB1 = OFF_XF + 3 * PS16;        // [64][64] head layer-1 activations (value, then action)
constexpr uint32_t OFF_DY64 = OFF_XB1 + 3 * PS64;      // [64][64] current upstream gradient
constexpr uint32_t OFF_DY16 = OFF_DY64 + 3 * PS64;     // [64][16]
// fp32 staging (no MMA operand).  Gathered row data of a tile (cp.async destinations), double-buffered so that the
// next tile's gather can be issued while this tile still reads its own: fp32 observations [64][8], then [5][64]
// 32-bit planes -- return, old value, action bits, old log-prob, advantage -- and the valid flags [64] u8; then the
// action layer-2 outputs [64][16] and d(loss)/d(head) [64][8] the mapping layer's gradients are accumulated from.
constexpr int R_RET = 0, R_VAL = 1, R_ACT = 2, R_LOGP = 3, R_ADV = 4, R_PLANES = 5;
constexpr uint32_t OBS32_BYTES = ROWS * 8 * 4, PLANES_BYTES = R_PLANES * ROWS * 4;
constexpr uint32_t OFF_OBS32 = OFF_DY16 + 3 * PS16;                  // [2][64][8] fp32
constexpr uint32_t OFF_PLANES = OFF_OBS32 + 2 * OBS32_BYTES;         // [2][5][64] 32-bit
constexpr uint32_t OFF_VALID = OFF_PLANES + 2 * PLANES_BYTES;        // [2][64] u8
constexpr uint32_t OFF_A2 = OFF_VALID + 2 * ROWS;                    // [64][16] fp32
constexpr uint32_t OFF_DHEAD = OFF_A2 + ROWS * 16 * 4;               // [64][8] fp32
constexpr uint32_t OFF_F32 = OFF_DHEAD + ROWS * 8 * 4;  // fp32 biases etc. (tc::F_COUNT floats)
constexpr uint32_t OFF_VPART = OFF_F32 + tc::F_RESERVE * 4;      // [4 column groups][64 rows] partial value dot products
constexpr uint32_t OFF_MISC = OFF_VPART + 4 * ROWS * 4;  // 2 mbarriers, TMEM slot
constexpr uint32_t OFF_IDX = OFF_MISC + 64;               // minibatch indices of the tile after the one being gathered
constexpr uint32_t SMEM_BYTES = OFF_IDX + ROWS * 4;
static_assert(SMEM_BYTES <= 227 * 1024, "fused update kernel exceeds shared memory");
static_assert(tc::F_COUNT <= tc::F_RESERVE, "fp32 parameter block too small");

// ---- TMEM columns (fp32; M = 64 accumulators use lanes 0-15 of every 32-lane quadrant)
constexpr uint32_t C_D0 = 0;       // 64: layer outputs
constexpr uint32_t C_D1 = 64;      // 64: action head layer-1 pre-activations (issued with the value head's)
constexpr uint32_t C_DF = 128;     // 16: d(features), summed over both heads
constexpr uint32_t C_G0 = 144;     // enc0  dW^T [64 out][16 in]
constexpr uint32_t C_G1 = 168;     // enc1  dW^T [64][64]
constexpr uint32_t C_G2 = 240;     // enc2  dW   [64 in][16 out]
constexpr uint32_t C_GV0 = 272;    // v0    dW^T [64][16]
constexpr uint32_t C_GV1 = 296;    // v1    dW^T [64][64]
constexpr uint32_t C_GA0 = 384;    // a0    dW^T [64][16]
constexpr uint32_t C_GA1 = 408;    // a1    dW   [64 in][16 out]
constexpr uint32_t TMEM_COLS = 512;

struct FusedArgs {
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
  int n_tiles;
  int obs_dim, n_out, is_box;
  tc::LayerOffsets lo;
};

__device__ __forceinline__ void tie_weights_f(float a, float b, float& wa, float& wb) {
  wa = a < b ? 1.f : (a == b ? 0.5f : 0.f);  // JAX: d min(a,b) splits ties 1/2-1/2
  wb = 1.f - wa;
}
__device__ __forceinline__ float clip_grad_f(float x, float lo, float hi) {
  return (x > lo && x < hi) ? 1.f : ((x == lo || x == hi) ? 0.5f : 0.f);
}

// NOUT: number of action-head outputs when known at compile time (1 = scalar Box, 2 = CartPole, 3 = Acrobot), so the
// per-row loss code of the row-owner threads -- the longest serial section of a tile -- unrolls without dead
// predicated iterations; 0 = read it from the arguments (up to LRX_MAX_NOUT).
template <int NOUT>
__global__ void __launch_bounds__(THREADS, 1) ppo_grad_fused_kernel(const __grid_constant__ FusedArgs a) {
  const int n_out = NOUT ? NOUT : a.n_out;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = umma::warp_idx_sync(), lane = tid & 31;
  const bool worker = warp < WORKERS / 32;  // warp-uniform
  const int q = warp & 3, g = worker ? (warp >> 2) : 0;
  const bool active = worker && lane < 16;
  const int m = q * 16 + (lane & 15);  // row of the tile owned by this thread (when active)
  const bool g0 = worker && (g == 0);
  float* fp = reinterpret_cast<float*>(smem + OFF_F32);
  float* vpart = reinterpret_cast<float*>(smem + OFF_VPART);
  uint64_t* barA = reinterpret_cast<uint64_t*>(smem + OFF_MISC);      // critical GEMM(s) of a step done
  uint64_t* barB = reinterpret_cast<uint64_t*>(smem + OFF_MISC + 8);  // background GEMMs of a step done
  uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + OFF_MISC + 16);
  uint64_t* barL = reinterpret_cast<uint64_t*>(smem + OFF_MISC + 24);  // every epilogue warp has done its TMEM loads
  int* idx_s = reinterpret_cast<int*>(smem + OFF_IDX);
  // current (read by this tile) and next (being gathered) staging buffers; swapped at the end of a tile
  float* obs_s = reinterpret_cast<float*>(smem + OFF_OBS32);
  float* obs_n = obs_s + OBS32_BYTES / 4;
  float* rows_s = reinterpret_cast<float*>(smem + OFF_PLANES);
  float* rows_n = rows_s + PLANES_BYTES / 4;
  uint8_t* valid_s = smem + OFF_VALID;
  uint8_t* valid_n = valid_s + ROWS;
  float* a2_s = reinterpret_cast<float*>(smem + OFF_A2);
  float* dh_s = reinterpret_cast<float*>(smem + OFF_DHEAD);

  const tc::Mat W1{smem + OFF_W1, 64, PS64}, W2{smem + OFF_W2, 16, PS16};
  const tc::Mat WV0{smem + OFF_WV0, 64, PS16}, WV1{smem + OFF_WV1, 64, PS64};
  const tc::Mat WA0{smem + OFF_WA0, 64, PS16}, WA1{smem + OFF_WA1, 16, PS16};
  // observations are double-buffered: the last GEMM of a tile (layer-0 weight gradient) still reads them while
  // the next tile's are written
  const tc::Mat XOBS2[2] = {tc::Mat{smem + OFF_XOBS, ROWS, PS16}, tc::Mat{smem + OFF_W0, ROWS, PS16}};
  const tc::Mat XH1{smem + OFF_XH1, ROWS, PS64}, XH2{smem + OFF_XH2, ROWS, PS64};
  const tc::Mat DH1 = XH2;  // d(loss)/d(h1) of a tile reuses the XH2 buffer (see the last two steps of a tile)
  const tc::Mat XF{smem + OFF_XF, ROWS, PS16}, XB1{smem + OFF_XB1, ROWS, PS64};
  const tc::Mat DY64{smem + OFF_DY64, ROWS, PS64}, DY16{smem + OFF_DY16, ROWS, PS16};

  // ------------------------------------------------------------------ one-time staging
  tc::stage_weight(a.params + a.lo.w[tc::L_ENC1], 64, 64, 64, W1, tid, THREADS);
  tc::stage_weight(a.params + a.lo.w[tc::L_ENC2], 16, 64, 64, W2, tid, THREADS);
  tc::stage_weight(a.params + a.lo.w[tc::L_V0], 64, 16, 16, WV0, tid, THREADS);
  tc::stage_weight(a.params + a.lo.w[tc::L_V1], 64, 64, 64, WV1, tid, THREADS);
  tc::stage_weight(a.params + a.lo.w[tc::L_A0], 64, 16, 16, WA0, tid, THREADS);
  tc::stage_weight(a.params + a.lo.w[tc::L_A1], 16, 64, 64, WA1, tid, THREADS);
  tc::stage_f32(a.params, a.lo, n_out, a.obs_dim, fp, tid, THREADS);
  if (tid == 0) {
    umma::mbar_init(barA, 1);
    umma::mbar_init(barB, 1);
    umma::mbar_init(barL, WORKERS / 32);
    umma::fence_mbar_init();
  }
  if (warp == 0) umma::tmem_alloc(tslot, TMEM_COLS);
  umma::fence_async_smem();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tb = *tslot;
  const uint32_t tw = umma::tmem_addr(tb, q * 32, 0);  // this warp's quadrant
  if (worker) {
    float z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0.f;
    for (int c = g * 128; c < g * 128 + 128; c += 16) umma::tmem_st16(tw + c, z);
    umma::tmem_st_wait();
  }

  // One step = the epilogue warps publish their shared-memory operands, the issuer warp enqueues the
  // step's CRITICAL GEMM(s) (the result the next epilogue consumes) and commits barA, then the
  // BACKGROUND GEMMs (weight / bias gradients accumulating in TMEM) and commits barB.  An epilogue
  // waits for barA, does its TMEM load and arithmetic while the background GEMMs run, and waits
  // for barB only before overwriting a buffer they read.  Every phase of both barriers is consumed
  // before the next step's barrier, so a waiter is never more than one phase behind.
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
  // track_bg: commit the background GEMMs to barB.  Only needed where something must know that THEY are done
  // before the next critical GEMM is: an epilogue that overwrites one of their operands (need_bg), and the last
  // step of a tile.  Everywhere else the tensor pipe's in-order execution is enough -- the next critical GEMM's
  // commit implies theirs -- and not waiting for barB saves its ~300-cycle commit latency per step.
  // after_loads: a tcgen05.ld does not overtake MMAs that are already queued (tools/tmem_ld_under_mma.py: it
  // returns when the queue has all but drained), so background GEMMs enqueued right behind the critical one delay
  // the epilogue by their full length.  The issuer then holds them back until every epilogue warp has signalled
  // (barL) that its TMEM loads of this step are done; they run under the epilogue's arithmetic and stores.
  uint32_t phL = 0;
  bool pendL = false;
  auto signal_loaded = [&]() {  // epilogue warps, after tcgen05.wait::ld (or with nothing to load)
    if (pendL) {
      pendL = false;
      umma::tc_fence_before();
      if (lane == 0) umma::mbar_arrive(barL);
    }
  };
  auto step_impl = [&](auto&& crit, auto&& bg, bool track_bg, bool after_loads = false) {
    if (worker) {
      signal_loaded();
      wait_crit();
      wait_bg();
      umma::fence_async_smem();
    }
    umma::tc_fence_before();
    __syncthreads();
    if (!worker) {
      umma::tc_fence_after();
      const bool leader = umma::elect_one();
      crit(leader);
      if (leader) umma::mma_commit(barA);
      if (after_loads) {
        umma::mbar_wait(barL, phL);
        phL ^= 1u;
        umma::tc_fence_after();
      }
      bg(leader);
      if (track_bg && leader) umma::mma_commit(barB);
      __syncwarp();
    } else {
      pendA = true;
      pendB = track_bg;
      pendL = after_loads;
    }
  };
  auto step_late = [&](auto&& crit, auto&& bg) { step_impl(crit, bg, false, true); };
  auto step = [&](auto&& crit, auto&& bg) { step_impl(crit, bg, false); };
  auto step_track = [&](auto&& crit, auto&& bg) { step_impl(crit, bg, true); };
  auto none = [](bool) {};

  const LrxPpoHyper h = a.h;
  const float invB = (float)(1.0 / a.B_global);
  float adv_mean = 0.f, adv_den = 1.f;
  if (h.normalize_advantages) {
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
  double s_kl = 0, s_pol = 0, s_val = 0, s_ent = 0, s_dls = 0, s_bad = 0;

  const int cg = g * 16;  // first column of this thread's group in a 64-wide layer

  // 64-wide layers use the 16x256b TMEM load shape so that ALL 32 lanes of the warp hold a fragment of
  // the quadrant's 16 rows x this group's 16 columns: rows fr, fr+8 and column pairs fc, fc+8.
  const int fr0 = q * 16 + (lane >> 2), fr1 = fr0 + 8;  // tile rows of this thread's fragment
  const int fc = cg + (lane & 3) * 2;                   // columns fc, fc+1, fc+8, fc+9
  // fragment order: v[0,1] = (fr0, fc..), v[2,3] = (fr1, fc..), v[4,5] = (fr0, fc+8..), v[6,7] = (fr1, fc+8..)
  auto frag_store = [&](const tc::Mat& dst, const float* v, bool need_bg) {
    uint32_t w[12];
    tc::split2(v[0], v[1], w[0], w[1], w[2]);
    tc::split2(v[2], v[3], w[3], w[4], w[5]);
    tc::split2(v[4], v[5], w[6], w[7], w[8]);
    tc::split2(v[6], v[7], w[9], w[10], w[11]);
    if (need_bg) wait_bg();
    tc::store2(dst, fr0, fc, w[0], w[1], w[2]);
    tc::store2(dst, fr1, fc, w[3], w[4], w[5]);
    tc::store2(dst, fr0, fc + 8, w[6], w[7], w[8]);
    tc::store2(dst, fr1, fc + 8, w[9], w[10], w[11]);
  };
  // layer output -> + bias -> ReLU (8-bit mask recorded) -> operand X; fn(v) sees the fragment
  auto frag_fwd = [&](uint32_t col, const float* bias, const tc::Mat& dst, uint32_t& mask, auto&& fn) {
    wait_crit();
    float v[8];
    umma::tmem_ld_frag16(tw + col + cg, v);
    umma::tmem_ld_wait();
    signal_loaded();
    const float2 b0 = *reinterpret_cast<const float2*>(bias + fc), b1 = *reinterpret_cast<const float2*>(bias + fc + 8);
    v[0] += b0.x; v[1] += b0.y; v[2] += b0.x; v[3] += b0.y;
    v[4] += b1.x; v[5] += b1.y; v[6] += b1.x; v[7] += b1.y;
    mask = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (v[j] > 0.f) mask |= (1u << j);
      v[j] = fmaxf(v[j], 0.f);
    }
    fn(v);
    frag_store(dst, v, false);
  };
  // the same without the operand store: the activated fragment stays in registers (value head layer 2, whose only
  // consumers -- the 64 -> 1 output layer and its weight gradient -- run on the FMA pipe)
  auto frag_fwd_keep = [&](uint32_t col, const float* bias, uint32_t& mask, float* v) {
    wait_crit();
    umma::tmem_ld_frag16(tw + col + cg, v);
    umma::tmem_ld_wait();
    const float2 b0 = *reinterpret_cast<const float2*>(bias + fc), b1 = *reinterpret_cast<const float2*>(bias + fc + 8);
    v[0] += b0.x; v[1] += b0.y; v[2] += b0.x; v[3] += b0.y;
    v[4] += b1.x; v[5] += b1.y; v[6] += b1.x; v[7] += b1.y;
    mask = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (v[j] > 0.f) mask |= (1u << j);
      v[j] = fmaxf(v[j], 0.f);
    }
  };
  // dX fragment masked by the ReLU of the layer that produced X -> operand dY
  // Bias gradients of the five 64-wide layers: every thread sums the dY values of its fragment (two rows x four
  // columns) over all tiles in registers; the cross-thread reduction happens once, at the flush.  As 12 extra
  // MMAs per layer and tile (dY^T x ones) they only lengthened the queue the epilogues' TMEM loads wait behind.
  float bs_e0[4] = {0.f, 0.f, 0.f, 0.f}, bs_e1[4] = {0.f, 0.f, 0.f, 0.f}, bs_v0[4] = {0.f, 0.f, 0.f, 0.f};
  float bs_v1[4] = {0.f, 0.f, 0.f, 0.f}, bs_a0[4] = {0.f, 0.f, 0.f, 0.f};
  // value output layer (64 -> 1): weight gradient of this thread's four columns and (one thread per row) the bias
  float gw_v2[4] = {0.f, 0.f, 0.f, 0.f}, gb_v2 = 0.f;
  // action_dist.mapping (16 -> n_out): this thread's share of dW[o][i] and db[o], o = 2q + lane/16, i = lane%16,
  // over the 16 rows of the tile that belong to its column group
  float gm_w = 0.f, gm_b = 0.f;
  float bs_e2[4] = {0.f, 0.f, 0.f, 0.f};  // encoder output layer's bias: columns 4g..4g+3 of this (row-owner) thread
  auto bias_add = [](float* bsum, const float* v) {
    bsum[0] += v[0] + v[2];
    bsum[1] += v[1] + v[3];
    bsum[2] += v[4] + v[6];
    bsum[3] += v[5] + v[7];
  };
  auto frag_bwd = [&](uint32_t col, uint32_t mask, const tc::Mat& dst, bool need_bg, float* bsum) {
    wait_crit();
    float v[8];
    umma::tmem_ld_frag16(tw + col + cg, v);
    umma::tmem_ld_wait();
    signal_loaded();
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = ((mask >> j) & 1u) ? v[j] : 0.f;
    bias_add(bsum, v);
    frag_store(dst, v, need_bg);
  };
  auto nopf = [](float*) {};
  // 16-wide layer outputs (features, d(features)): every column group takes 4 of the 16 columns of its
  // quadrant's rows (lane = row), so the step is not left to one group while twelve warps wait.
  auto quarter16 = [&](uint32_t col, const float* bias, const tc::Mat& dst, float* qsum) {
    wait_crit();
    float v[4];
    umma::tmem_ld4(tw + col + 4 * g, v);
    umma::tmem_ld_wait();
    signal_loaded();
    if (bias) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += bias[4 * g + j];
    }
    if (qsum && active) {  // column sums over the tiles (bias gradient), reduced across threads at the flush
#pragma unroll
      for (int j = 0; j < 4; ++j) qsum[j] += v[j];
    }
    uint32_t w[6];
    tc::split2(v[0], v[1], w[0], w[1], w[2]);
    tc::split2(v[2], v[3], w[3], w[4], w[5]);
    if (active) {
      uint8_t* d = dst.ptr + (uint32_t)((4 * g) >> 3) * (dst.R * 16u) + (uint32_t)m * 16u + (uint32_t)((4 * g) & 7) * 2u;
      *reinterpret_cast<uint2*>(d) = make_uint2(w[0], w[3]);
      *reinterpret_cast<uint2*>(d + dst.piece_bytes) = make_uint2(w[1], w[4]);
      *reinterpret_cast<uint2*>(d + 2 * dst.piece_bytes) = make_uint2(w[2], w[5]);
    }
  };

  // Row data of a tile (gather of base_buffer.py:90-103): cp.async straight into shared memory, issued by the
  // otherwise idle lanes 16-31 of column group 2 (one lane per row) one tile ahead -- while the row-owner threads of
  // group 0 work through the policy loss and groups 2 and 3 have nothing to do --, the minibatch index two tiles
  // ahead.  No register is the destination of a global load inside the tile loop: a register gather -- even one
  // consumed a whole tile later -- cost 9 % of the kernel, because everything that shares a scoreboard with the
  // outstanding loads waits out their DRAM latency.  The staging is double-buffered (see the shared-memory map).
  auto cp_async4 = [](void* dst, const void* src, bool ok) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(umma::smem_u32(dst)), "l"(src), "r"(ok ? 4 : 0) : "memory");
  };
  auto cp_async16 = [](void* dst, const void* src, bool ok) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(umma::smem_u32(dst)), "l"(src), "r"(ok ? 16 : 0) : "memory");
  };
  auto tile_row_ok = [&](int t) { return t < a.n_tiles && (long long)t * ROWS + m < a.B; };
  auto gather_rows = [&](int i, bool ok) {  // buffer row i (zeros when !ok) -> this lane's staging slot m
    const int n = i / a.buf.T, t = i - n * a.buf.T;
    const size_t src = ok ? (size_t)t * a.buf.N + n : 0;
    cp_async4(rows_n + R_RET * ROWS + m, a.buf.ret + src, ok);
    cp_async4(rows_n + R_VAL * ROWS + m, (h.clip_value_loss ? a.buf.val : a.buf.ret) + src, ok && h.clip_value_loss);
    cp_async4(rows_n + R_ACT * ROWS + m, (const int32_t*)a.buf.act + src, ok);
    cp_async4(rows_n + R_LOGP * ROWS + m, a.buf.logp + src, ok);
    cp_async4(rows_n + R_ADV * ROWS + m, a.buf.adv + src, ok);  // raw; normalised where it is used
    const float* op = a.buf.obs + src * a.obs_dim;
    if (a.obs_dim == 4) {  // one 16-byte copy (CartPole)
      cp_async16(obs_n + m * 8, op, ok);
      *reinterpret_cast<float4*>(obs_n + m * 8 + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) cp_async4(obs_n + m * 8 + k, op + (k < a.obs_dim ? k : 0), ok && k < a.obs_dim);
    }
    valid_n[m] = ok ? 1 : 0;
  };
  auto gather_idx = [&](int t) {  // minibatch indices of tile t -> idx_s
    const bool ok = tile_row_ok(t);
    cp_async4(idx_s + m, a.idx + (ok ? (long long)t * ROWS + m : 0), ok);
  };
  auto gather_commit = []() { asm volatile("cp.async.commit_group;" ::: "memory"); };
  auto gather_wait = []() { asm volatile("cp.async.wait_group 0;" ::: "memory"); };
  const bool gatherer = worker && g == 2 && !active;
  auto swap_staging = [&]() {
    float* t = obs_s; obs_s = obs_n; obs_n = t;
    t = rows_s; rows_s = rows_n; rows_n = t;
    uint8_t* u = valid_s; valid_s = valid_n; valid_n = u;
  };
  if (gatherer) {  // first tile: the only register load of an index
    const int t0 = (int)blockIdx.x;
    const bool ok = tile_row_ok(t0);
    const int i = ok ? __ldg(a.idx + (long long)t0 * ROWS + m) : 0;
    gather_rows(i, ok);
    gather_idx(t0 + (int)gridDim.x);
    gather_commit();
    gather_wait();
  }
  swap_staging();  // what was gathered as "next" is the first tile's "current"
  __syncthreads();

  auto store_xobs = [&](const tc::Mat& xo) {
    float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, o[8];
    const float4 lo4 = *reinterpret_cast<const float4*>(obs_s + m * 8), hi4 = *reinterpret_cast<const float4*>(obs_s + m * 8 + 4);
    o[0] = lo4.x; o[1] = lo4.y; o[2] = lo4.z; o[3] = lo4.w; o[4] = hi4.x; o[5] = hi4.y; o[6] = hi4.z; o[7] = hi4.w;
    tc::store8(xo, 0, m, o);
    tc::store8(xo, 1, m, z8);
  };
  int tile_parity = 0;
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const tc::Mat& XOBS = XOBS2[tile_parity];
    tile_parity ^= 1;
    if (g0 && active) store_xobs(XOBS);
    const bool valid = g0 && active && valid_s[m] != 0;  // this thread owns a real row of the minibatch

    // ---------------------------------------------------------------- encoder forward
    uint32_t mk_h1 = 0, mk_h2 = 0, mk_b1 = 0, mk_b2 = 0;
    // encoder layer 0 (d <= 8 inputs) on the FMA pipe, straight from the gathered observation: no MMA round
    // trip.  The fragment rows' observations come from the fp32 staging written by the gathering lanes.
    if (worker) {
      const int l0 = lane >> 2, l1 = l0 + 8;
      float v[8];
      {
        const float2 b0 = *reinterpret_cast<const float2*>(fp + tc::F_B0 + fc);
        const float2 b1 = *reinterpret_cast<const float2*>(fp + tc::F_B0 + fc + 8);
        v[0] = b0.x; v[1] = b0.y; v[2] = b0.x; v[3] = b0.y; v[4] = b1.x; v[5] = b1.y; v[6] = b1.x; v[7] = b1.y;
      }
      const float* w0p = fp + tc::F_W0;
      float xa[8], xb[8];
      {
        const float4* p0 = reinterpret_cast<const float4*>(obs_s + (q * 16 + l0) * 8);
        const float4* p1 = reinterpret_cast<const float4*>(obs_s + (q * 16 + l1) * 8);
        const float4 a0 = p0[0], b0 = p1[0];
        xa[0] = a0.x; xa[1] = a0.y; xa[2] = a0.z; xa[3] = a0.w;
        xb[0] = b0.x; xb[1] = b0.y; xb[2] = b0.z; xb[3] = b0.w;
        if (a.obs_dim > 4) {
          const float4 a1 = p0[1], b1 = p1[1];
          xa[4] = a1.x; xa[5] = a1.y; xa[6] = a1.z; xa[7] = a1.w;
          xb[4] = b1.x; xb[5] = b1.y; xb[6] = b1.z; xb[7] = b1.w;
        } else {
          xa[4] = xa[5] = xa[6] = xa[7] = xb[4] = xb[5] = xb[6] = xb[7] = 0.f;
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (k < a.obs_dim) {
          const float x0 = xa[k], x1 = xb[k];
          const float wa = w0p[fc * 8 + k], wb = w0p[(fc + 1) * 8 + k], wc = w0p[(fc + 8) * 8 + k], wd = w0p[(fc + 9) * 8 + k];
          v[0] = fmaf(wa, x0, v[0]); v[1] = fmaf(wb, x0, v[1]); v[2] = fmaf(wa, x1, v[2]); v[3] = fmaf(wb, x1, v[3]);
          v[4] = fmaf(wc, x0, v[4]); v[5] = fmaf(wd, x0, v[5]); v[6] = fmaf(wc, x1, v[6]); v[7] = fmaf(wd, x1, v[7]);
        }
      }
      mk_h1 = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (v[j] > 0.f) mk_h1 |= (1u << j);
        v[j] = fmaxf(v[j], 0.f);
      }
      frag_store(XH1, v, true);  // the previous tile's last GEMMs read XH1 / XH2
    }
    step([&](bool ld) { tc::gemm<false, false>(ld, tb + C_D0, XH1, W1, 64, 64, 4, 0u); }, none);
    if (worker) frag_fwd(C_D0, fp + tc::F_B1, XH2, mk_h2, nopf);
    step([&](bool ld) { tc::gemm<false, false>(ld, tb + C_D0, XH2, W2, 64, 16, 4, 0u); }, none);
    if (worker) quarter16(C_D0, fp + tc::F_B2, XF, nullptr);

    // ---------------------------------------------------------------- value head forward
    step([&](bool ld) { tc::gemm<false, false>(ld, tb + C_D0, XF, WV0, 64, 64, 1, 0u); },
         [&](bool ld) { tc::gemm<false, false>(ld, tb + C_D1, XF, WA0, 64, 64, 1, 0u); });  // action head layer 1, consumed later
    if (worker) frag_fwd(C_D0, fp + tc::F_BV0, XB1, mk_b1, nopf);
    step([&](bool ld) { tc::gemm<false, false>(ld, tb + C_D0, XB1, WV1, 64, 64, 4, 0u); }, none);
    float xb2[8];  // value layer-2 activations of this thread's fragment
    if (worker) {
      float pr0, pr1;  // partial value-head dot products of rows fr0, fr1 over this thread's 4 columns
      frag_fwd_keep(C_D0, fp + tc::F_BV1, mk_b2, xb2);
      {
        const float2 w0 = *reinterpret_cast<const float2*>(fp + tc::F_WV2 + fc);
        const float2 w1 = *reinterpret_cast<const float2*>(fp + tc::F_WV2 + fc + 8);
        pr0 = fmaf(w1.y, xb2[5], fmaf(w1.x, xb2[4], fmaf(w0.y, xb2[1], w0.x * xb2[0])));
        pr1 = fmaf(w1.y, xb2[7], fmaf(w1.x, xb2[6], fmaf(w0.y, xb2[3], w0.x * xb2[2])));
      }
      pr0 += __shfl_xor_sync(0xffffffffu, pr0, 1);
      pr1 += __shfl_xor_sync(0xffffffffu, pr1, 1);
      pr0 += __shfl_xor_sync(0xffffffffu, pr0, 2);
      pr1 += __shfl_xor_sync(0xffffffffu, pr1, 2);
      if ((lane & 3) == 0) {
        vpart[g * ROWS + fr0] = pr0;
        vpart[g * ROWS + fr1] = pr1;
      }
    }
    __syncthreads();
    auto value_of = [&](int row) {
      return (((fp[tc::F_BV2] + vpart[row]) + vpart[ROWS + row]) + vpart[2 * ROWS + row]) + vpart[3 * ROWS + row];
    };
    const float value = active ? value_of(m) : 0.f;

    // ---------------------------------------------------------------- value loss (ppo.py:435-451)
    // d loss / d value of one row (0 for rows past the minibatch) and its squared-error statistic
    auto value_grad = [&](float val, float ret_r, float vold_r, bool valid_r, float& sv) -> float {
      sv = 0.f;
      if (!valid_r) return 0.f;
      const float c = h.clip_coefficient;
      const float e1 = val - ret_r;
      float dvl;
      if (h.clip_value_loss) {
        const float dv = val - vold_r;
        const float clipped = vold_r + lrx_clipf(dv, -c, c);
        const float e2 = clipped - ret_r;
        const float q1 = e1 * e1, q2 = e2 * e2;
        sv = fminf(q1, q2);
        float u1, u2;
        tie_weights_f(q1, q2, u1, u2);
        dvl = invB * (u1 * e1 + u2 * e2 * clip_grad_f(dv, -c, c));
      } else {
        sv = e1 * e1;
        dvl = invB * e1;
      }
      return h.value_loss_coefficient * dvl;
    };
    bool bad = false;
    if (worker) {
      // the fragment rows' scalars come from the staging written by the gathering lanes
      const float ret0 = rows_s[R_RET * ROWS + fr0], ret1 = rows_s[R_RET * ROWS + fr1];
      const float vo0 = rows_s[R_VAL * ROWS + fr0], vo1 = rows_s[R_VAL * ROWS + fr1];
      const bool va0 = valid_s[fr0] != 0, va1 = valid_s[fr1] != 0;
      float sv_unused;
      const float dv0 = value_grad(value_of(fr0), ret0, vo0, va0, sv_unused);
      const float dv1 = value_grad(value_of(fr1), ret1, vo1, va1, sv_unused);
      // ---------------------------------------------------------------- value head backward (64 -> 1 on the FMA pipe)
      gw_v2[0] = fmaf(xb2[0], dv0, fmaf(xb2[2], dv1, gw_v2[0]));
      gw_v2[1] = fmaf(xb2[1], dv0, fmaf(xb2[3], dv1, gw_v2[1]));
      gw_v2[2] = fmaf(xb2[4], dv0, fmaf(xb2[6], dv1, gw_v2[2]));
      gw_v2[3] = fmaf(xb2[5], dv0, fmaf(xb2[7], dv1, gw_v2[3]));
      if (g0 && (lane & 3) == 0) gb_v2 += dv0 + dv1;
      const float2 w0 = *reinterpret_cast<const float2*>(fp + tc::F_WV2 + fc);
      const float2 w1 = *reinterpret_cast<const float2*>(fp + tc::F_WV2 + fc + 8);
      float v[8] = {dv0 * w0.x, dv0 * w0.y, dv1 * w0.x, dv1 * w0.y, dv0 * w1.x, dv0 * w1.y, dv1 * w1.x, dv1 * w1.y};
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = ((mk_b2 >> j) & 1u) ? v[j] : 0.f;
      bias_add(bs_v1, v);
      frag_store(DY64, v, false);
      if (g0 && active) {
        float sv;
        value_grad(value, rows_s[R_RET * ROWS + m], rows_s[R_VAL * ROWS + m], valid, sv);
        if (valid) {
          bad = !isfinite(value);
          s_val += (double)sv;
        }
      }
    }
    step_track([&](bool ld) { tc::gemm<false, true>(ld, tb + C_D0, DY64, WV1, 64, 64, 4, 0u); },        // dX = dY W
         [&](bool ld) {
           tc::gemm<true, true>(ld, tb + C_GV1, DY64, XB1, 64, 64, 4, 1u);                        // dW^T[out][in] += dY^T X
         });
    if (worker) {
      frag_bwd(C_D0, mk_b1, DY64, true, bs_v0);  // background GEMMs read DY64 and XB1
      // -------------------------------------------------------------- action head forward
      frag_fwd(C_D1, fp + tc::F_BA0, XB1, mk_b1, nopf);
    }
    // background, issued once the action head's output has been loaded from TMEM: d(features) from the value head
    // and the value layer-0 weight gradient run under the policy-loss arithmetic (the longest stretch of a tile
    // without tensor work).  Nothing touches DY64 / XF / C_DF / C_GV0 before the next critical GEMM is done.
    step_late([&](bool ld) { tc::gemm<false, false>(ld, tb + C_D0, XB1, WA1, 64, 16, 4, 0u); },
              [&](bool ld) {
                tc::gemm<false, true>(ld, tb + C_DF, DY64, WV0, 64, 16, 4, 0u);
                tc::gemm<true, true>(ld, tb + C_GV0, DY64, XF, 64, 16, 4, 1u);
              });
    if (gatherer) {  // the next tile's rows, into the staging buffers this tile does not read
      gather_rows(idx_s[m], tile_row_ok(tile + (int)gridDim.x));
      gather_idx(tile + 2 * (int)gridDim.x);
      gather_commit();
    }
    if (g0) {
      float a2[16];
      wait_crit();
      umma::tmem_ld16(tw + C_D0, a2);
      umma::tmem_ld_wait();
      signal_loaded();
#pragma unroll
      for (int j = 0; j < 16; ++j) a2[j] += fp[tc::F_BA1 + j];
      if (active) {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(a2_s + m * 16 + j) = make_float4(a2[j], a2[j + 1], a2[j + 2], a2[j + 3]);
      }
      // action_dist.mapping (Linear 16 -> n_out, policy/actor.py:248-261) on the FMA pipe
      float head[LRX_MAX_NOUT];
#pragma unroll
      for (int o2 = 0; o2 < LRX_MAX_NOUT; ++o2) {
        float sacc = fp[tc::F_BMAP + o2];
        if (o2 < n_out) {
#pragma unroll
          for (int i = 0; i < 16; ++i) sacc = fmaf(fp[tc::F_WMAP + o2 * 16 + i], a2[i], sacc);
        }
        head[o2] = sacc;
      }

      // -------------------------------------------------------------- policy loss (ppo.py:405-434,452-467)
      float dhead[LRX_MAX_NOUT];
#pragma unroll
      for (int o2 = 0; o2 < LRX_MAX_NOUT; ++o2) dhead[o2] = 0.f;
      if (valid) {
        const int act_raw = reinterpret_cast<const int*>(rows_s)[R_ACT * ROWS + m];  // fp32 (Box) or int32 (Discrete) bits
        const float logp_old = rows_s[R_LOGP * ROWS + m], A = rows_s[R_ADV * ROWS + m];
        float logp, entropy, z_n = 0.f;
        float lp[LRX_MAX_NOUT], pr[LRX_MAX_NOUT];
        if (a.is_box) {
          const float loc = head[0];
          z_n = (__int_as_float(act_raw) - loc) / scale;
          logp = -0.5f * (z_n * z_n) - (LRX_HALF_LOG_2PI + log_scale);
          entropy = 0.5f + LRX_HALF_LOG_2PI + log_scale;
        } else {
#pragma unroll
          for (int j = 0; j < LRX_MAX_NOUT; ++j) lp[j] = j < n_out ? head[j] : 0.f;
          log_softmax_n<LRX_MAX_NOUT>(lp, n_out);
          logp = 0.f;
          entropy = 0.f;
#pragma unroll
          for (int j = 0; j < LRX_MAX_NOUT; ++j) {
            if (j < n_out) {
              pr[j] = expf(lp[j]);
              entropy -= pr[j] * lp[j];
              if (j == act_raw) logp = lp[j];
            }
          }
        }
        bad = bad || !(isfinite(logp) && isfinite(entropy));  // ppo.py:413-417
        if (bad) s_bad += 1.0;
        const float log_ratio = logp - logp_old;
        const float ratio = expf(log_ratio);
        s_kl += (double)(ratio - log_ratio);
        const float c = h.clip_coefficient, lo = 1.f - c, hi = 1.f + c;
        const float An = h.normalize_advantages ? (A - adv_mean) / adv_den : A;  // ppo.py:424-427
        float dlogp;
        if (h.surrogate == 1) {  // A2C / REINFORCE: -mean(logp * A)   (a2c.py:401, reinforce.py:391)
          s_pol += (double)(logp * An);
          dlogp = -invB * An;
        } else {
          const float s1 = An * ratio, s2 = An * lrx_clipf(ratio, lo, hi);
          s_pol += (double)fminf(s1, s2);
          float w1, w2;
          tie_weights_f(s1, s2, w1, w2);
          dlogp = -invB * (w1 * An * ratio + w2 * An * clip_grad_f(ratio, lo, hi) * ratio);
        }
        s_ent += (double)entropy;
        const float ch = h.entropy_loss_coefficient;
        if (a.is_box) {
          dhead[0] = dlogp * z_n / scale;
          s_dls += (double)(dlogp * (z_n * z_n - 1.0f));
        } else {
#pragma unroll
          for (int j = 0; j < LRX_MAX_NOUT; ++j) {
            if (j < n_out) {
              const float dH = -pr[j] * (lp[j] + entropy);
              dhead[j] = dlogp * ((j == act_raw ? 1.f : 0.f) - pr[j]) + (-ch * invB) * dH;
            }
          }
        }
      }
      // -------------------------------------------------------------- action head backward (mapping, FMA)
      if (active) {
        *reinterpret_cast<float4*>(dh_s + m * 8) = make_float4(dhead[0], dhead[1], dhead[2], dhead[3]);
        *reinterpret_cast<float4*>(dh_s + m * 8 + 4) = make_float4(dhead[4], dhead[5], dhead[6], dhead[7]);
        float d16[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float sacc = 0.f;
#pragma unroll
          for (int o2 = 0; o2 < LRX_MAX_NOUT; ++o2)
            if (o2 < n_out) sacc = fmaf(dhead[o2], fp[tc::F_WMAP + o2 * 16 + i], sacc);
          d16[i] = sacc;
        }
        tc::store8(DY16, 0, m, d16);
        tc::store8(DY16, 1, m, d16 + 8);
      }
    }
    step([&](bool ld) { tc::gemm<false, true>(ld, tb + C_D0, DY16, WA1, 64, 64, 1, 0u); },        // dX = dY16 W_a1 (K = 16 outputs)
         [&](bool ld) {
           tc::gemm<true, true>(ld, tb + C_GA1, XB1, DY16, 64, 16, 4, 1u);                        // dW[in][out] += X^T dY
         });
    // mapping layer gradients on the FMA pipe while the GEMM above runs: dW[o][i] += dhead[r][o] a2[r][i],
    // db[o] += dhead[r][o] over this column group's 16 rows (the a1 bias gradient follows from db at the flush)
    if (worker) {
      const int o2 = 2 * q + (lane >> 4), i2 = lane & 15;
      if (o2 < n_out) {
        const float* ap = a2_s + (g * 16) * 16 + i2;
        const float* dp = dh_s + (g * 16) * 8 + o2;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const float dh = dp[r * 8];
          gm_w = fmaf(ap[r * 16], dh, gm_w);
          gm_b += dh;
        }
      }
    }
    if (worker) frag_bwd(C_D0, mk_b1, DY64, false, bs_a0);  // this step's background GEMMs do not read DY64
    step([&](bool ld) { tc::gemm<false, true>(ld, tb + C_DF, DY64, WA0, 64, 16, 4, 1u); },        // d(features) += action head
         [&](bool ld) {
           tc::gemm<true, true>(ld, tb + C_GA0, DY64, XF, 64, 16, 4, 1u);
         });

    // ---------------------------------------------------------------- encoder backward
    if (worker) quarter16(C_DF, nullptr, DY16, bs_e2);                   // background reads DY64 / XF only
    step([&](bool ld) { tc::gemm<false, true>(ld, tb + C_D0, DY16, W2, 64, 64, 1, 0u); },
         [&](bool ld) {
           tc::gemm<true, true>(ld, tb + C_G2, XH2, DY16, 64, 16, 4, 1u);
         });
    if (worker) frag_bwd(C_D0, mk_h2, DY64, false, bs_e1);  // background reads XH2 / DY16 only
    step([&](bool ld) { tc::gemm<false, true>(ld, tb + C_D0, DY64, W1, 64, 64, 4, 0u); },
         [&](bool ld) {
           tc::gemm<true, true>(ld, tb + C_G1, DY64, XH1, 64, 64, 4, 1u);
         });
    // d(h1) goes into the XH2 buffer, dead since the GEMMs of the step before (ordered before this step's critical
    // GEMM) read it: no wait for this step's weight-gradient GEMM, which still reads DY64 and XH1
    if (worker) frag_bwd(C_D0, mk_h1, DH1, false, bs_e0);
    // the next tile's rows have had three steps to arrive; the barrier of the step below publishes them
    if (gatherer) gather_wait();
    step_track(none, [&](bool ld) {
      tc::gemm<true, true>(ld, tb + C_G0, DH1, XOBS, 64, 16, 4, 1u);
    });  // its commit also covers the GEMM on XH1 of the step before: waited for before XH1 / XH2 are rewritten
    swap_staging();

  }
  if (worker) {
    wait_crit();
    wait_bg();
  }

  // ------------------------------------------------------------------ flush the gradient partial
  float* part = a.partial + (size_t)blockIdx.x * a.partial_stride;
  const tc::LayerOffsets& lo = a.lo;
  // dW^T accumulators: TMEM row = output unit (this thread's m), columns = inputs (the bias comes from registers)
  auto flush_T = [&](uint32_t col, int in_real, int in_pad, int layer) {
    for (int c0 = 0; c0 < in_pad; c0 += 8) {
      float v[8];
      umma::tmem_ld8(tw + col + c0, v);
      umma::tmem_ld_wait();
      if (active) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c0 + j < in_real) part[lo.w[layer] + m * in_real + c0 + j] = v[j];
      }
    }
  };
  // dW accumulators: TMEM row = input unit (m), columns = outputs (the biases come from registers)
  auto flush_N = [&](uint32_t col, int in_real, int out_real, int out_pad, int layer) {
    for (int c0 = 0; c0 < out_pad; c0 += 8) {
      float v[8];
      umma::tmem_ld8(tw + col + c0, v);
      umma::tmem_ld_wait();
      if (active) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c0 + j < out_real && m < in_real) part[lo.w[layer] + (c0 + j) * in_real + m] = v[j];
      }
    }
  };
  // nine layers over the four column groups of warps
  umma::tc_fence_after();
  if (!worker) { /* issuer warp: nothing to flush */ } else
  if (g == 0) { flush_T(C_G0, a.obs_dim, 16, tc::L_ENC0); flush_N(C_G2, 64, 16, 16, tc::L_ENC2); }
  if (worker && g == 1) flush_T(C_G1, 64, 64, tc::L_ENC1);
  if (worker && g == 2) { flush_T(C_GV1, 64, 64, tc::L_V1); flush_T(C_GV0, 16, 16, tc::L_V0); }
  if (worker && g == 3) { flush_T(C_GA0, 16, 16, tc::L_A0); flush_N(C_GA1, 64, 16, 16, tc::L_A1); }

  // ------------------------------------------------------------------ register bias sums (fixed order)
  __syncthreads();  // every GEMM is done and waited for: XH1 is free
  {
    float* bsh = reinterpret_cast<float*>(smem + OFF_XH1);  // [6 vectors][4 quadrants][64 columns], then [4] db_v2
    float* const bsv[6] = {bs_e0, bs_e1, bs_v0, bs_v1, bs_a0, gw_v2};
    if (worker) {
      {
        float x = gb_v2;  // lanes 0, 4, .., 28 of column group 0 hold the quadrant's rows
        x += __shfl_xor_sync(0xffffffffu, x, 4);
        x += __shfl_xor_sync(0xffffffffu, x, 8);
        x += __shfl_xor_sync(0xffffffffu, x, 16);
        if (g0 && lane == 0) bsh[6 * 4 * 64 + q] = x;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {  // encoder output bias: the quadrant's 16 row-owner lanes (the others hold 0)
        float x = bs_e2[j];
        x += __shfl_xor_sync(0xffffffffu, x, 1);
        x += __shfl_xor_sync(0xffffffffu, x, 2);
        x += __shfl_xor_sync(0xffffffffu, x, 4);
        x += __shfl_xor_sync(0xffffffffu, x, 8);
        x += __shfl_xor_sync(0xffffffffu, x, 16);
        if (lane == 0) bsh[6 * 4 * 64 + 4 + q * 16 + 4 * g + j] = x;
      }
#pragma unroll
      for (int l = 0; l < 6; ++l) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float x = bsv[l][j];  // rows lane/4 and lane/4 + 8 of the quadrant -> all 16 rows
          x += __shfl_xor_sync(0xffffffffu, x, 4);
          x += __shfl_xor_sync(0xffffffffu, x, 8);
          x += __shfl_xor_sync(0xffffffffu, x, 16);
          if (lane < 4) bsh[(l * 4 + q) * 64 + fc + (j & 1) + (j >> 1) * 8] = x;
        }
      }
    }
    __syncthreads();
    const int bl[6] = {lo.b[tc::L_ENC0], lo.b[tc::L_ENC1], lo.b[tc::L_V0], lo.b[tc::L_V1], lo.b[tc::L_A0], lo.w[tc::L_V2]};
    {  // mapping layer: sum the four row groups; db_a1 = W_map^T db_map (the layers are linear in between)
      float* mw = bsh + 6 * 4 * 64 + 4 + 64;  // [4 groups][128 (o, i)] then [4][8] db
      float* mb = mw + 4 * 128;
      if (worker) {
        const int p2 = q * 32 + lane;
        mw[g * 128 + p2] = gm_w;
        if ((lane & 15) == 0) mb[g * 8 + (p2 >> 4)] = gm_b;
      }
      __syncthreads();
      if (tid < 128 && (tid >> 4) < n_out)
        part[lo.w[tc::L_MAP] + tid] = ((mw[tid] + mw[128 + tid]) + mw[256 + tid]) + mw[384 + tid];
      if (tid >= 128 && tid < 136 && tid - 128 < n_out)
        part[lo.b[tc::L_MAP] + tid - 128] = ((mb[tid - 128] + mb[8 + tid - 128]) + mb[16 + tid - 128]) + mb[24 + tid - 128];
      if (tid >= 160 && tid < 176) {
        const int i2 = tid - 160;
        float sacc = 0.f;
        for (int o2 = 0; o2 < n_out; ++o2)
          sacc = fmaf(((mb[o2] + mb[8 + o2]) + mb[16 + o2]) + mb[24 + o2], fp[tc::F_WMAP + o2 * 16 + i2], sacc);
        part[lo.b[tc::L_A1] + i2] = sacc;
      }
    }
    if (tid >= 32 && tid < 48) {
      const float* e2 = bsh + 6 * 4 * 64 + 4 + (tid - 32);
      part[lo.b[tc::L_ENC2] + tid - 32] = ((e2[0] + e2[16]) + e2[32]) + e2[48];
    }
    if (tid == 0) part[lo.b[tc::L_V2]] = ((bsh[6 * 4 * 64] + bsh[6 * 4 * 64 + 1]) + bsh[6 * 4 * 64 + 2]) + bsh[6 * 4 * 64 + 3];
    for (int e = tid; e < 6 * 64; e += THREADS) {
      const int l = e >> 6, c = e & 63;
      part[bl[l] + c] = ((bsh[(l * 4 + 0) * 64 + c] + bsh[(l * 4 + 1) * 64 + c]) + bsh[(l * 4 + 2) * 64 + c]) + bsh[(l * 4 + 3) * 64 + c];
    }
  }

  // ------------------------------------------------------------------ loss statistics (fixed order)
  __syncthreads();
  double* sh = reinterpret_cast<double*>(smem + OFF_DY64);  // [6][4 warps]
  double v6[6] = {s_kl, s_pol, s_val, s_ent, s_dls, s_bad};
  if (g0) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      v6[k] = warp_sum(v6[k]);
      if (lane == 0) sh[k * 4 + q] = v6[k];
    }
  }
  __syncthreads();
  if (tid < 6) {
    double s = 0;
    for (int w = 0; w < 4; ++w) s += sh[tid * 4 + w];
    a.loss_partial[(size_t)blockIdx.x * 8 + tid] = s;
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
FusedWorkspace carve_fused(const LrxPolicyDesc& pol, void* base) {
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

size_t lrx_fused_workspace_bytes(const LrxPolicyDesc* pol, int64_t B) {
  if (!tc::is_default_policy(*pol)) return 0;
  return carve_fused(*pol, nullptr).total;
}


int lrx_ppo_grad_fused(const LrxPolicyDesc* pol, const float* params, const LrxBatchView* buf,
                       const int32_t* idx, int64_t B, int64_t B_global, const double* adv_sums,
                       const LrxPpoHyper* h, float* gradbuf, void* workspace, size_t workspace_bytes,
                       cudaStream_t stream, bool* handled, const LrxApplyArgs* apply) {
  *handled = false;
  if (!tc::is_default_policy(*pol) || !lrx_fused_enabled()) return LRX_OK;
  FusedWorkspace w = carve_fused(*pol, workspace);
  LRX_REQUIRE(workspace_bytes >= w.total, LRX_ERR_INVALID_ARG, "lrx_ppo_grad: workspace too small for the fused path");
  FusedArgs a{};
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
  a.n_tiles = (int)((B + ROWS - 1) / ROWS);
  a.obs_dim = pol->obs_dim;
  a.n_out = pol->n_out;
  a.is_box = pol->is_box;
  a.lo = tc::layer_offsets(*pol);
  const int sms = lrx_sm_count();
  const int grid = a.n_tiles < sms ? a.n_tiles : sms;
  static bool set0[64], set1[64], set2[64], set3[64];
  LRX_CUDA_CHECK(lrx_set_max_smem(ppo_grad_fused_kernel<0>, (int)SMEM_BYTES, set0));
  LRX_CUDA_CHECK(lrx_set_max_smem(ppo_grad_fused_kernel<1>, (int)SMEM_BYTES, set1));
  LRX_CUDA_CHECK(lrx_set_max_smem(ppo_grad_fused_kernel<2>, (int)SMEM_BYTES, set2));
  LRX_CUDA_CHECK(lrx_set_max_smem(ppo_grad_fused_kernel<3>, (int)SMEM_BYTES, set3));
  switch (pol->n_out) {
    case 1: ppo_grad_fused_kernel<1><<<grid, THREADS, SMEM_BYTES, stream>>>(a); break;
    case 2: ppo_grad_fused_kernel<2><<<grid, THREADS, SMEM_BYTES, stream>>>(a); break;
    case 3: ppo_grad_fused_kernel<3><<<grid, THREADS, SMEM_BYTES, stream>>>(a); break;
    default: ppo_grad_fused_kernel<0><<<grid, THREADS, SMEM_BYTES, stream>>>(a); break;
  }
  LRX_LAUNCH_CHECK();
  const int rc = lrx_launch_reduce_partials(w.partial, grid, (long long)w.P_pad, pol->n_params, w.loss_partial, grid,
                                            a.lo.log_std, h->entropy_loss_coefficient, (double)B_global, gradbuf, apply, stream);
  if (rc) return rc;
  LRX_LAUNCH_CHECK();
  *handled = true;
  return LRX_OK;
}
