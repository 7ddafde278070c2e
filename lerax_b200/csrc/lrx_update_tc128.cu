// Fused PPO minibatch gradient for the default MLPActorCriticPolicy widths on the tensor cores, second generation:
// 128-row tiles (UMMA M = 128), fp32 emulated with TWO bf16 pieces (lrx_tc2.cuh), accumulators in TMEM.
//
// Reference: algorithm/ppo.py:396-469 (ppo_loss + filter_value_and_grad) over the gathered minibatch of
// buffer/base_buffer.py:90-103 -- the same contract as lrx_update_fused.cu (first generation: 64-row tiles, three
// pieces), which stays selectable with lrx_debug_set_fused(1) and is checked against this kernel in the tests.
//
// Why this shape.  The first-generation kernel was bound by a serial chain per tile -- critical MMA group -> commit
// -> tcgen05.ld -> epilogue -> CTA barrier, eleven times -- with one 64-row tile in flight and six bf16 MMAs per fp32
// product (profiles/r1_ppo_grad_fused.md).  The update's contract is 1e-4 (BASELINE.json north_star), not the
// rollout's 1e-5, so two round-to-nearest pieces and three products are enough (~3e-6 rms per dot product): half the
// MMAs and two thirds of the shared memory, which buys a 128-row tile.  The chain is walked once per 128 rows instead
// of once per 64, an M = 128 MMA costs 48 clk against 2 x 32 for two M = 64 ones, and every epilogue thread owns one
// ROW (TMEM lane = row, 16 consecutive columns: 32x32b loads, 16-byte operand stores).
// Bias gradients ride on the weight-gradient GEMMs: the activation operand carries a constant ones column, so
// dW^T = dY^T [X | 1] has the bias gradient as its last accumulator column (no separate reduction, no extra MMA).
//
// One persistent CTA per SM: 16 epilogue warps (warp w: TMEM quadrant w & 3 = rows 32q .. 32q+31, column group
// w >> 2 = 16 of a layer's 64 columns) + 1 issuer warp.  Per tile:
//   gather rows (cp.async) -> encoder layer 0 (FMA) -> enc1 -> enc2 -> value head -> value loss -> value head bwd ->
//   action head -> policy loss -> action head bwd -> encoder bwd,
// weight gradients accumulated in TMEM across all tiles of the CTA and flushed once as a per-CTA partial
// (reduce_partials_kernel sums the partials in a fixed order and applies clip + Adam).
#include <stdlib.h>

#include "lrx_common.cuh"
#include "lrx_policy.cuh"
#include "lrx_tc2.cuh"

namespace {

constexpr int ROWS = 128;              // rows per tile = UMMA M
constexpr int WORKERS = 512;           // 16 epilogue warps
constexpr int THREADS = WORKERS + 32;  // + warp 16: issues every tcgen05.mma (converged, one elected lane)

// ---- shared memory map (bytes).  A c-group (8 columns) of an R-row matrix is R * 16 bytes.
constexpr uint32_t CG = ROWS * 16;  // 2048: c-group of a 128-row activation matrix
constexpr uint32_t OFF_W1 = 0;                       // [64][64]  hi 8192 | lo 8192
constexpr uint32_t OFF_W2 = OFF_W1 + 16384;          // [16][64]  hi 2048 | lo 2048
constexpr uint32_t OFF_WV0 = OFF_W2 + 4096;          // [64][16]  hi 2048 | lo 2048
constexpr uint32_t OFF_WV1 = OFF_WV0 + 4096;         // [64][64]
constexpr uint32_t OFF_WA0 = OFF_WV1 + 16384;        // [64][16]
constexpr uint32_t OFF_WA1 = OFF_WA0 + 4096;         // [16][64]
constexpr uint32_t OFF_XOBS = OFF_WA1 + 4096;        // [128][8 obs | ones, 0 x 7]: hi 2 c-groups, lo 1 c-group
constexpr uint32_t OFF_XH1 = OFF_XOBS + 3 * CG;      // [128][64 | ones]: hi 9 c-groups, lo 8
constexpr uint32_t OFF_XH2 = OFF_XH1 + 17 * CG;      // [128][64]: hi 8, lo 8 (also d(loss)/d(h1) at the end of a tile)
constexpr uint32_t OFF_XF = OFF_XH2 + 16 * CG;       // [128][16 | ones]: hi 3, lo 2
constexpr uint32_t OFF_XB1 = OFF_XF + 5 * CG;        // [128][64 | ones]: head layer-1 activations (value, then action)
constexpr uint32_t OFF_DY64 = OFF_XB1 + 17 * CG;     // [128][64]: current upstream gradient
constexpr uint32_t OFF_DY16 = OFF_DY64 + 16 * CG;    // [128][16]
// fp32 staging (no MMA operand): gathered row data (cp.async destinations).  Observations [128][8]: single-buffered (a
// tile reads them only at its top, the gather of the next tile is issued later).  Five 32-bit planes [5][128] -- return,
// old value, action bits, old log-prob, advantage -- and the valid flags [128] u8: double-buffered, because the gather
// is issued under the policy loss, which still reads the current tile's.
constexpr int R_RET = 0, R_VAL = 1, R_ACT = 2, R_LOGP = 3, R_ADV = 4, R_PLANES = 5;
constexpr uint32_t OFF_OBS32 = OFF_DY16 + 4 * CG;
constexpr uint32_t OFF_PLANES = OFF_OBS32 + ROWS * 8 * 4;
constexpr uint32_t PLANES_BYTES = R_PLANES * ROWS * 4;
constexpr uint32_t OFF_VALID = OFF_PLANES + 2 * PLANES_BYTES;
constexpr uint32_t OFF_F32 = OFF_VALID + 2 * ROWS;                    // fp32 biases etc. (tc::F_COUNT floats)
constexpr uint32_t OFF_VPART = OFF_F32 + tc::F_RESERVE * 4;       // [4 column groups][128 rows] partial value dot products
constexpr uint32_t OFF_MISC = OFF_VPART + 4 * ROWS * 4;           // 3 mbarriers, TMEM slot
constexpr uint32_t OFF_IDX = OFF_MISC + 64;                       // minibatch indices of the tile after the one being gathered
constexpr uint32_t SMEM_BYTES = OFF_IDX + ROWS * 4;
static_assert(SMEM_BYTES <= 227 * 1024, "tc128 update kernel exceeds shared memory");
static_assert(tc::F_COUNT <= tc::F_RESERVE, "fp32 parameter block too small");

// ---- TMEM columns (fp32).  M = 128 accumulators use all 128 lanes; the M = 64 weight-gradient accumulators use
// lanes 0-15 of every 32-lane quadrant (output unit j at lane 32 * (j / 16) + j % 16).
constexpr uint32_t C_D0 = 0;       // 64: layer outputs
constexpr uint32_t C_D1 = 64;      // 64: action head layer-1 pre-activations (issued with the value head's)
constexpr uint32_t C_DF = 128;     // 16: d(features), summed over both heads
constexpr uint32_t C_G0 = 144;     // enc0  dW^T [64 out][8 in | bias | 0 x 7]
constexpr uint32_t C_G1 = 160;     // enc1  dW^T [64][64 | bias, 0 x 7]
constexpr uint32_t C_G2 = 232;     // enc2  dW   [64 in][16 out]
constexpr uint32_t C_GV0 = 248;    // v0    dW^T [64][16 | bias, 0 x 7]
constexpr uint32_t C_GV1 = 272;    // v1    dW^T [64][64 | bias, 0 x 7]
constexpr uint32_t C_GA0 = 344;    // a0    dW^T [64][16 | bias, 0 x 7]
constexpr uint32_t C_GA1 = 368;    // a1    dW   [64 in][16 out]
constexpr uint32_t TMEM_COLS = 512;

struct Tc128Args {
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
  long long* prof;  // test hook (lrx_debug_set_tc128_prof): per-step cycle totals of CTA 0, thread 0 and the issuer
};

__device__ __forceinline__ void tie_weights_f(float a, float b, float& wa, float& wb) {
  wa = a < b ? 1.f : (a == b ? 0.5f : 0.f);  // JAX: d min(a,b) splits ties 1/2-1/2
  wb = 1.f - wa;
}
__device__ __forceinline__ float clip_grad_f(float x, float lo, float hi) {
  return (x > lo && x < hi) ? 1.f : ((x == lo || x == hi) ? 0.5f : 0.f);
}

// Sum over the 32 lanes of 16 per-lane values, scattered: afterwards v[0] of lane L is the warp-wide sum of element
// (L >> 1) & 15 (both lanes of a pair hold the same sum).  16 shuffles instead of 80 for sixteen full reductions.
__device__ __forceinline__ float reduce_scatter16(float* v, int lane) {
#pragma unroll
  for (int half = 8, off = 16; half >= 1; half >>= 1, off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < half) {
        const float send = up ? v[j] : v[j + half];
        const float keep = up ? v[j + half] : v[j];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// NOUT: number of action-head outputs when known at compile time (1 = scalar Box, 2 = CartPole, 3 = Acrobot /
// MountainCar); 0 = read it from the arguments (up to LRX_MAX_NOUT).
template <int NOUT>
__global__ void __launch_bounds__(THREADS, 1) ppo_grad_tc128_kernel(const __grid_constant__ Tc128Args a) {
  constexpr int NO = NOUT ? NOUT : LRX_MAX_NOUT;
  const int n_out = NOUT ? NOUT : a.n_out;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = umma::warp_idx_sync(), lane = tid & 31;
  const bool worker = warp < WORKERS / 32;  // warp-uniform
  const int q = warp & 3, g = worker ? (warp >> 2) : 0;
  const int m = q * 32 + lane;  // row of the tile owned by this thread
  const bool g0 = worker && (g == 0);  // row-owner warps: per-row loss arithmetic
  const int cg = g * 16;               // first column of this thread's group in a 64-wide layer
  float* fp = reinterpret_cast<float*>(smem + OFF_F32);
  float* vpart = reinterpret_cast<float*>(smem + OFF_VPART);
  uint64_t* barA = reinterpret_cast<uint64_t*>(smem + OFF_MISC);      // critical GEMM(s) of a step done
  uint64_t* barB = reinterpret_cast<uint64_t*>(smem + OFF_MISC + 8);  // background GEMMs of a step done
  uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + OFF_MISC + 16);
  uint64_t* barL = reinterpret_cast<uint64_t*>(smem + OFF_MISC + 24);  // every epilogue warp has done its TMEM loads
  int* idx_s = reinterpret_cast<int*>(smem + OFF_IDX);
  float* obs_s = reinterpret_cast<float*>(smem + OFF_OBS32);
  // current (read by this tile) and next (being gathered) staging buffers; swapped at the end of a tile
  float* rows_s = reinterpret_cast<float*>(smem + OFF_PLANES);
  float* rows_n = rows_s + PLANES_BYTES / 4;
  uint8_t* valid_s = smem + OFF_VALID;
  uint8_t* valid_n = valid_s + ROWS;

  const tc2::Mat W1{smem + OFF_W1, 64, 8192}, W2{smem + OFF_W2, 16, 2048};
  const tc2::Mat WV0{smem + OFF_WV0, 64, 2048}, WV1{smem + OFF_WV1, 64, 8192};
  const tc2::Mat WA0{smem + OFF_WA0, 64, 2048}, WA1{smem + OFF_WA1, 16, 2048};
  const tc2::Mat XOBS{smem + OFF_XOBS, ROWS, 2 * CG};
  const tc2::Mat XH1{smem + OFF_XH1, ROWS, 9 * CG}, XH2{smem + OFF_XH2, ROWS, 8 * CG};
  const tc2::Mat DH1 = XH2;  // d(loss)/d(h1) of a tile reuses the XH2 buffer (see the last two steps of a tile)
  const tc2::Mat XF{smem + OFF_XF, ROWS, 3 * CG}, XB1{smem + OFF_XB1, ROWS, 9 * CG};
  const tc2::Mat DY64{smem + OFF_DY64, ROWS, 8 * CG}, DY16{smem + OFF_DY16, ROWS, 2 * CG};

  // ------------------------------------------------------------------ one-time staging
  {
    // The six tensor-core weight matrices are 1536 chunks of 8 consecutive inputs of one output unit.  Every thread issues
    // the loads of ALL its chunks (<= 3) before converting any: one L2 round trip for the whole staging instead of one per
    // matrix (six dependent round trips were 8 k cycles of every launch).
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
    tc::stage_f32(a.params, a.lo, n_out, a.obs_dim, fp, tid, THREADS);  // its loads join the ones in flight
#pragma unroll
    for (int k = 0; k < PER; ++k)
      if (mi[k] >= 0) tc2::store8(tc2::Mat{smem + base[mi[k]], (uint32_t)outs[mi[k]], lo_off[mi[k]]}, cc[k], oo[k], v[k]);
  }
  if (tid < ROWS) {  // the constant ones columns (bf16 1.0 = 0x3F80 in element 0 of the extra c-group)
    const uint4 one = make_uint4(0x00003F80u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(XOBS.ptr + 1 * CG + tid * 16) = one;
    *reinterpret_cast<uint4*>(XH1.ptr + 8 * CG + tid * 16) = one;
    *reinterpret_cast<uint4*>(XF.ptr + 2 * CG + tid * 16) = one;
    *reinterpret_cast<uint4*>(XB1.ptr + 8 * CG + tid * 16) = one;
  }
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

  // profiling marks (only with the debug hook set): slot k accumulates the cycles between mark k-1 and mark k of row-owner
  // thread 0 (slots 0..31) and of the issuer's lane 0 (slots 32..63)
  long long prof_last = 0;
  int prof_step = 1;
  const bool prof_on = a.prof != nullptr && blockIdx.x == 0 && (tid == 0 || tid == WORKERS);
  auto mark = [&](int k) {
    if (prof_on) {
      const long long t = clock64();
      if (k > 0) a.prof[(tid == 0 ? 0 : 32) + k] += t - prof_last;
      prof_last = t;
    }
  };

  // One step = the epilogue warps publish their shared-memory operands, the issuer warp enqueues the step's CRITICAL
  // GEMM(s) (the result the next epilogue consumes) and commits barA, then the BACKGROUND GEMMs (weight gradients
  // accumulating in TMEM) and commits barB where something must know that they are done.  See lrx_update_fused.cu for
  // the measurements behind track_bg / after_loads (tcgen05.ld does not overtake queued MMAs; a commit costs ~300 clk).
  uint32_t phA = 0, phB = 0, phL = 0;
  bool pendA = false, pendB = false, pendL = false;
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
    long long tbar = 0;
    if (prof_on) tbar = clock64();
    __syncthreads();
    if (prof_on) {  // slots 17..31 (thread 0) / 49..63 (issuer): cycles spent waiting at the barrier of step 1, 2, ..
      a.prof[(tid == 0 ? 16 : 48) + prof_step] += clock64() - tbar;
      ++prof_step;
    }
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
  // per-thread loss statistics: one row per tile, at most a few dozen tiles per launch -> fp32 here, fp64 across threads
  float s_kl = 0, s_pol = 0, s_val = 0, s_ent = 0, s_dls = 0, s_bad = 0;

  // ---- epilogue building blocks: this thread's row m, columns cg .. cg+15 of a 64-wide layer
  auto store_row16 = [&](const tc2::Mat& dst, const float* v, bool need_bg) {
    uint32_t hw[8], lw[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) tc2::split2(v[2 * i], v[2 * i + 1], hw[i], lw[i]);
    if (need_bg) wait_bg();
    uint8_t* d = dst.ptr + (uint32_t)(2 * g) * CG + (uint32_t)m * 16u;
    *reinterpret_cast<uint4*>(d) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(d + CG) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
    *reinterpret_cast<uint4*>(d + dst.lo_off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    *reinterpret_cast<uint4*>(d + dst.lo_off + CG) = make_uint4(lw[4], lw[5], lw[6], lw[7]);
  };
  auto add_bias_relu = [&](float* v, const float* bias, uint32_t& mask) {
    const float4* bp = reinterpret_cast<const float4*>(bias + cg);  // warp-uniform address: broadcast
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 b = bp[i];
      v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
    }
    // ReLU + its mask in two instructions per element: the sign bits are funnel-shifted into `mask` (element j ends up at
    // bit 15 - j, 1 = negative pre-activation = no gradient).  A pre-activation of exactly +0 passes its (zero-measure)
    // gradient where jax.nn.relu would not.
    mask = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      mask = __funnelshift_l(__float_as_uint(v[j]), mask, 1);
      v[j] = fmaxf(v[j], 0.f);
    }
  };
  // layer output -> + bias -> ReLU (16-bit mask recorded) -> operand X
  auto fwd64 = [&](uint32_t col, const float* bias, const tc2::Mat& dst, uint32_t& mask) {
    wait_crit();
    float v[16];
    umma::tmem_ld16(tw + col + cg, v);
    umma::tmem_ld_wait();
    signal_loaded();
    add_bias_relu(v, bias, mask);
    store_row16(dst, v, false);
  };
  // dX masked by the ReLU of the layer that produced X -> operand dY
  auto bwd64 = [&](uint32_t col, uint32_t mask, const tc2::Mat& dst, bool need_bg) {
    wait_crit();
    float v[16];
    umma::tmem_ld16(tw + col + cg, v);
    umma::tmem_ld_wait();
    signal_loaded();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = ((mask >> (15 - j)) & 1u) ? 0.f : v[j];
    store_row16(dst, v, need_bg);
  };
  // 16-wide layer outputs (features, d(features)): every column group takes 4 of the 16 columns of its rows
  float bs_e2[4] = {0.f, 0.f, 0.f, 0.f};  // encoder output layer's bias gradient: columns 4g..4g+3 over this thread's rows
  auto quarter16 = [&](uint32_t col, const float* bias, const tc2::Mat& dst, bool sum_cols) {
    wait_crit();
    float v[4];
    umma::tmem_ld4(tw + col + 4 * g, v);
    umma::tmem_ld_wait();
    signal_loaded();
    if (bias) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += bias[4 * g + j];
    }
    if (sum_cols) {
#pragma unroll
      for (int j = 0; j < 4; ++j) bs_e2[j] += v[j];
    }
    tc2::store4(dst, 4 * g, m, v);
  };

  // ---- row gather (buffer/base_buffer.py:90-103): cp.async straight into shared staging, one tile ahead; the
  // minibatch index two tiles ahead.  No register is the destination of a global load in
  // the tile loop (lrx_update_fused.cu: such a load cost 9 %).
  auto cp_async4 = [](void* dst, const void* src, bool ok) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(umma::smem_u32(dst)), "l"(src), "r"(ok ? 4 : 0) : "memory");
  };
  auto cp_async16 = [](void* dst, const void* src, bool ok) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(umma::smem_u32(dst)), "l"(src), "r"(ok ? 16 : 0) : "memory");
  };
  auto tile_row_ok = [&](int t, int r) { return t < a.n_tiles && (long long)t * ROWS + r < a.B; };
  auto gather_row = [&](int r, int i, bool ok) {  // buffer row i (zeros when !ok) -> staging slot r
    if (a.buf.packed != nullptr) {
      // row-major record (lrx_ppo_pack_rows): one 64-byte line per row instead of one sector per plane
      const float* rec = a.buf.packed + (ok ? (size_t)i * LRX_PACKED_FLOATS : 0);
      cp_async16(obs_s + r * 8, rec, ok);
      cp_async16(obs_s + r * 8 + 4, rec + 4, ok);
      cp_async4(rows_n + R_ACT * ROWS + r, rec + 8, ok);
      cp_async4(rows_n + R_LOGP * ROWS + r, rec + 9, ok);
      cp_async4(rows_n + R_VAL * ROWS + r, rec + 10, ok && h.clip_value_loss);
      cp_async4(rows_n + R_RET * ROWS + r, rec + 11, ok);
      cp_async4(rows_n + R_ADV * ROWS + r, rec + 12, ok);  // raw; normalised where it is used
      valid_n[r] = ok ? 1 : 0;
      return;
    }
    const int n = i / a.buf.T, t = i - n * a.buf.T;
    const size_t src = ok ? (size_t)t * a.buf.N + n : 0;
    cp_async4(rows_n + R_RET * ROWS + r, a.buf.ret + src, ok);
    cp_async4(rows_n + R_VAL * ROWS + r, (h.clip_value_loss ? a.buf.val : a.buf.ret) + src, ok && h.clip_value_loss);
    cp_async4(rows_n + R_ACT * ROWS + r, (const int32_t*)a.buf.act + src, ok);
    cp_async4(rows_n + R_LOGP * ROWS + r, a.buf.logp + src, ok);
    cp_async4(rows_n + R_ADV * ROWS + r, a.buf.adv + src, ok);  // raw; normalised where it is used
    const float* op = a.buf.obs + src * a.obs_dim;
    if (a.obs_dim == 4) {  // one 16-byte copy (CartPole)
      cp_async16(obs_s + r * 8, op, ok);
      *reinterpret_cast<float4*>(obs_s + r * 8 + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) cp_async4(obs_s + r * 8 + k, op + (k < a.obs_dim ? k : 0), ok && k < a.obs_dim);
    }
    valid_n[r] = ok ? 1 : 0;
  };
  // Column group 2's warps gather, one lane per row, right after the barrier of the policy-loss step, during which
  // they have nothing else to do: the DRAM latency of the copies runs under the row-owner warps' loss arithmetic.  (The
  // first fence.proxy.async any warp of the CTA executes after the copies were issued waits for them to land -- its
  // MEMBAR drains the CTA's outstanding LDGSTS -- so they must be issued early in a long step, not late: measured with
  // tools/tc128_prof.py, 1.7 k cycles per tile.)
  auto gather_tile = [&](int t_next, int t_after, bool first) {
    const bool ok = tile_row_ok(t_next, m);
    const int i = first ? (ok ? __ldg(a.idx + (long long)t_next * ROWS + m) : 0) : idx_s[m];
    gather_row(m, i, ok);
    const bool ok2 = tile_row_ok(t_after, m);  // minibatch index of the tile after: read only by this lane, next time
    cp_async4(idx_s + m, a.idx + (ok2 ? (long long)t_after * ROWS + m : 0), ok2);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto gather_wait = []() { asm volatile("cp.async.wait_group 0;" ::: "memory"); };
  auto swap_staging = [&]() {
    float* t = rows_s; rows_s = rows_n; rows_n = t;
    uint8_t* u = valid_s; valid_s = valid_n; valid_n = u;
  };
  const bool gatherer = worker && g == 2;
  if (gatherer) {  // first tile: the only register loads of indices
    gather_tile((int)blockIdx.x, (int)blockIdx.x + (int)gridDim.x, true);
    gather_wait();
  }
  swap_staging();  // what was gathered as "next" is the first tile's "current"
  __syncthreads();

  // gradients kept in registers across the tiles of the CTA
  float gw_v2 = 0.f, gb_v2 = 0.f;  // value output layer (64 -> 1): dW of column cg + (lane >> 1) % 16 over this warp's rows; db
  // action_dist.mapping (16 -> n_out): dW[o][4 (lane % 4) + c] over this lane's rows (lane / 4 of every tile's octet), db[o]
  float gm_w[NO][4], gm_b[NO];
#pragma unroll
  for (int o2 = 0; o2 < NO; ++o2) gm_w[o2][0] = gm_w[o2][1] = gm_w[o2][2] = gm_w[o2][3] = gm_b[o2] = 0.f;

  mark(0);
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    mark(14);
    prof_step = 1;
    const bool valid_m = valid_s[m] != 0;  // this thread's row is a real row of the minibatch
    mark(0);

    // ---------------------------------------------------------------- encoder layer 0 (d <= 8 inputs) on the FMA pipe
    uint32_t mk_h1 = 0, mk_h2 = 0, mk_b1 = 0, mk_b2 = 0;
    if (worker) {
      float x[8];
      {
        const float4 lo4 = *reinterpret_cast<const float4*>(obs_s + m * 8), hi4 = *reinterpret_cast<const float4*>(obs_s + m * 8 + 4);
        x[0] = lo4.x; x[1] = lo4.y; x[2] = lo4.z; x[3] = lo4.w; x[4] = hi4.x; x[5] = hi4.y; x[6] = hi4.z; x[7] = hi4.w;
      }
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 wa = *reinterpret_cast<const float4*>(fp + tc::F_W0 + (cg + j) * 8);  // warp-uniform: broadcast
        float acc = fmaf(wa.w, x[3], fmaf(wa.z, x[2], fmaf(wa.y, x[1], wa.x * x[0])));
        if (a.obs_dim > 4) {
          const float4 wb = *reinterpret_cast<const float4*>(fp + tc::F_W0 + (cg + j) * 8 + 4);
          acc = fmaf(wb.w, x[7], fmaf(wb.z, x[6], fmaf(wb.y, x[5], fmaf(wb.x, x[4], acc))));
        }
        v[j] = acc;
      }
      add_bias_relu(v, fp + tc::F_B0, mk_h1);
      store_row16(XH1, v, true);  // waits for the previous tile's last GEMMs (they read XH1 / XH2 / XOBS)
      if (g0) {
        float z8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) z8[k] = k < a.obs_dim ? x[k] : 0.f;
        tc2::store8(XOBS, 0, m, z8);  // lo piece: c-group 0 only (XOBS.lo_off = 2 c-groups)
      }
    }
    mark(1);
    step([&](bool ld) { tc2::gemm<false, false>(ld, tb + C_D0, XH1, W1, ROWS, 64, 64, 4, 0u); }, none);
    if (worker) fwd64(C_D0, fp + tc::F_B1, XH2, mk_h2);
    mark(2);
    step([&](bool ld) { tc2::gemm<false, false>(ld, tb + C_D0, XH2, W2, ROWS, 16, 16, 4, 0u); }, none);
    if (worker) quarter16(C_D0, fp + tc::F_B2, XF, false);

    mark(3);
    // ---------------------------------------------------------------- value head forward
    step([&](bool ld) { tc2::gemm<false, false>(ld, tb + C_D0, XF, WV0, ROWS, 64, 64, 1, 0u); },
         [&](bool ld) { tc2::gemm<false, false>(ld, tb + C_D1, XF, WA0, ROWS, 64, 64, 1, 0u); });  // action head layer 1, consumed later
    if (worker) fwd64(C_D0, fp + tc::F_BV0, XB1, mk_b1);
    mark(4);
    step([&](bool ld) { tc2::gemm<false, false>(ld, tb + C_D0, XB1, WV1, ROWS, 64, 64, 4, 0u); }, none);
    float xb2[16];  // value layer-2 activations of this thread's row and column group
    if (worker) {
      wait_crit();
      umma::tmem_ld16(tw + C_D0 + cg, xb2);
      umma::tmem_ld_wait();
      add_bias_relu(xb2, fp + tc::F_BV1, mk_b2);
      const float4* wp = reinterpret_cast<const float4*>(fp + tc::F_WV2 + cg);
      float pr = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 w = wp[i];
        pr = fmaf(w.w, xb2[4 * i + 3], fmaf(w.z, xb2[4 * i + 2], fmaf(w.y, xb2[4 * i + 1], fmaf(w.x, xb2[4 * i], pr))));
      }
      vpart[g * ROWS + m] = pr;
    }
    __syncthreads();

    mark(5);
    // ---------------------------------------------------------------- value loss (ppo.py:435-451) and its backward
    if (worker) {
      const float value = (((fp[tc::F_BV2] + vpart[m]) + vpart[ROWS + m]) + vpart[2 * ROWS + m]) + vpart[3 * ROWS + m];
      const float ret_r = rows_s[R_RET * ROWS + m], vold_r = rows_s[R_VAL * ROWS + m];
      float dv = 0.f, sv = 0.f;  // d loss / d value of this row (0 past the minibatch) and its squared-error statistic
      if (valid_m) {
        const float c = h.clip_coefficient;
        const float e1 = value - ret_r;
        float dvl;
        if (h.clip_value_loss) {
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
        dv = h.value_loss_coefficient * dvl;
      }
      if (g0 && valid_m) {
        if (!isfinite(value)) s_bad += 1.0f;  // ppo.py:413-417 (a row with a non-finite log-prob too counts twice: only > 0 matters)
        s_val += sv;
        gb_v2 += dv;
      }
      // value output layer (64 -> 1) on the FMA pipe: dW[c] += h2[r][c] dv[r] (reduced over the warp's rows), dX = dv W
      float t16[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) t16[j] = xb2[j] * dv;
      gw_v2 += reduce_scatter16(t16, lane);
      const float4* wp = reinterpret_cast<const float4*>(fp + tc::F_WV2 + cg);
      float v[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 w = wp[i];
        v[4 * i] = dv * w.x; v[4 * i + 1] = dv * w.y; v[4 * i + 2] = dv * w.z; v[4 * i + 3] = dv * w.w;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = ((mk_b2 >> (15 - j)) & 1u) ? 0.f : v[j];
      store_row16(DY64, v, false);
    }
    mark(6);
    step_track([&](bool ld) { tc2::gemm<false, true>(ld, tb + C_D0, DY64, WV1, ROWS, 64, 64, 4, 0u); },      // dX = dY W
               [&](bool ld) { tc2::gemm<true, true>(ld, tb + C_GV1, DY64, XB1, 64, 72, 64, 8, 1u); });        // [dW^T | db] += dY^T [X | 1]
    if (worker) {
      bwd64(C_D0, mk_b1, DY64, true);  // the background GEMM reads DY64 and XB1
      // -------------------------------------------------------------- action head forward
      fwd64(C_D1, fp + tc::F_BA0, XB1, mk_b1);
    }
    mark(7);
    // background, issued once the action head's output has been loaded from TMEM: d(features) from the value head and
    // the value layer-0 weight gradient run under the policy-loss arithmetic
    step_late([&](bool ld) { tc2::gemm<false, false>(ld, tb + C_D0, XB1, WA1, ROWS, 16, 16, 4, 0u); },
              [&](bool ld) {
                tc2::gemm<false, true>(ld, tb + C_DF, DY64, WV0, ROWS, 16, 16, 4, 0u);
                tc2::gemm<true, true>(ld, tb + C_GV0, DY64, XF, 64, 24, 16, 8, 1u);
              });
    // the next tile's rows, fetched under the policy loss (the longest step): observations into the single staging
    // buffer (dead since the tile top), the per-row planes into the buffers this tile does not read
    if (gatherer) gather_tile(tile + (int)gridDim.x, tile + 2 * (int)gridDim.x, false);
    if (worker) {
      // ------------------------------------------------------------ action head output, policy loss and their backward,
      // on ALL sixteen warps: four lanes per row.  Left to the four row-owner warps (one per scheduler, no latency
      // hiding) this was the longest step of a tile -- ~800 dependent instructions, 4.5 k cycles (tools/tc128_prof.py).
      // Warp (q, g) serves rows 32q + 8g .. + 7; lane = 4 * (row % 8) + p holds features 4p .. 4p + 3 of its row.
      const int p4 = lane & 3, rl = 8 * g + (lane >> 2), rr = 32 * q + rl;
      float x[4];
      {
        float a2[16];
        wait_crit();
        umma::tmem_ld16(tw + C_D0, a2);  // lane l: row 32q + l, all 16 features
        umma::tmem_ld_wait();
        signal_loaded();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float v0 = __shfl_sync(0xffffffffu, a2[c], rl), v1 = __shfl_sync(0xffffffffu, a2[4 + c], rl);
          const float v2 = __shfl_sync(0xffffffffu, a2[8 + c], rl), v3 = __shfl_sync(0xffffffffu, a2[12 + c], rl);
          x[c] = p4 == 0 ? v0 : (p4 == 1 ? v1 : (p4 == 2 ? v2 : v3));
        }
        const float4 b4 = *reinterpret_cast<const float4*>(fp + tc::F_BA1 + 4 * p4);
        x[0] += b4.x; x[1] += b4.y; x[2] += b4.z; x[3] += b4.w;
      }
      // action_dist.mapping (Linear 16 -> n_out, policy/actor.py:248-261) on the FMA pipe: partial over this lane's four
      // features, then the row's four lanes are summed (all four end up with the logits / loc)
      float head[NO], wm[NO][4];
#pragma unroll
      for (int o2 = 0; o2 < NO; ++o2) {
        float sacc = 0.f;
        if (o2 < n_out) {
          const float4 w4 = *reinterpret_cast<const float4*>(fp + tc::F_WMAP + o2 * 16 + 4 * p4);
          wm[o2][0] = w4.x; wm[o2][1] = w4.y; wm[o2][2] = w4.z; wm[o2][3] = w4.w;
          sacc = fmaf(w4.w, x[3], fmaf(w4.z, x[2], fmaf(w4.y, x[1], w4.x * x[0])));
          sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
          sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
          sacc += fp[tc::F_BMAP + o2];
        } else {
          wm[o2][0] = wm[o2][1] = wm[o2][2] = wm[o2][3] = 0.f;
        }
        head[o2] = sacc;
      }

      // -------------------------------------------------------------- policy loss (ppo.py:405-434,452-467), computed by
      // the row's four lanes alike; lane p = 0 keeps the statistics
      float dhead[NO];
#pragma unroll
      for (int o2 = 0; o2 < NO; ++o2) dhead[o2] = 0.f;
      if (valid_s[rr] != 0) {
        const int act_raw = reinterpret_cast<const int*>(rows_s)[R_ACT * ROWS + rr];  // fp32 (Box) or int32 (Discrete) bits
        const float logp_old = rows_s[R_LOGP * ROWS + rr], A = rows_s[R_ADV * ROWS + rr];
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
        const float c = h.clip_coefficient, lo = 1.f - c, hi = 1.f + c;
        const float An = h.normalize_advantages ? (A - adv_mean) / adv_den : A;  // ppo.py:424-427
        float dlogp, spol;
        if (h.surrogate == 1) {  // A2C / REINFORCE: -mean(logp * A)   (a2c.py:401, reinforce.py:391)
          spol = logp * An;
          dlogp = -invB * An;
        } else {
          const float s1 = An * ratio, s2 = An * lrx_clipf(ratio, lo, hi);
          spol = fminf(s1, s2);
          float w1, w2;
          tie_weights_f(s1, s2, w1, w2);
          dlogp = -invB * (w1 * An * ratio + w2 * An * clip_grad_f(ratio, lo, hi) * ratio);
        }
        const float ch = h.entropy_loss_coefficient;
        float sdls = 0.f;
        if (a.is_box) {
          dhead[0] = dlogp * z_n / scale;
          sdls = dlogp * (z_n * z_n - 1.0f);
        } else {
#pragma unroll
          for (int j = 0; j < NO; ++j) {
            if (j < n_out) {
              const float dH = -pr[j] * (lp[j] + entropy);
              dhead[j] = dlogp * ((j == act_raw ? 1.f : 0.f) - pr[j]) + (-ch * invB) * dH;
            }
          }
        }
        if (p4 == 0) {
          if (bad_p) s_bad += 1.0f;
          s_kl += ratio - log_ratio;
          s_pol += spol;
          s_ent += entropy;
          s_dls += sdls;
#pragma unroll
          for (int o2 = 0; o2 < NO; ++o2) gm_b[o2] += dhead[o2];
        }
      }
      // -------------------------------------------------------------- action head backward (mapping, FMA): this lane's
      // four columns of d(a2), and its share of the mapping layer's weight gradient dW[o][4p + c] += dhead[o] a2[4p + c]
      {
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
        tc2::store4(DY16, 4 * p4, rr, d4);
      }
    }
    mark(8);
    step([&](bool ld) { tc2::gemm<false, true>(ld, tb + C_D0, DY16, WA1, ROWS, 64, 64, 1, 0u); },        // dX = dY16 W_a1 (K = 16 outputs)
         [&](bool ld) { tc2::gemm<true, true>(ld, tb + C_GA1, XB1, DY16, 64, 16, 16, 8, 1u); });           // dW[in][out] += X^T dY
    if (worker) bwd64(C_D0, mk_b1, DY64, false);  // this step's background GEMM does not read DY64
    mark(9);
    step([&](bool ld) { tc2::gemm<false, true>(ld, tb + C_DF, DY64, WA0, ROWS, 16, 16, 4, 1u); },        // d(features) += action head
         [&](bool ld) { tc2::gemm<true, true>(ld, tb + C_GA0, DY64, XF, 64, 24, 16, 8, 1u); });

    // ---------------------------------------------------------------- encoder backward
    if (worker) quarter16(C_DF, nullptr, DY16, true);                    // background reads DY64 / XF only
    mark(10);
    step([&](bool ld) { tc2::gemm<false, true>(ld, tb + C_D0, DY16, W2, ROWS, 64, 64, 1, 0u); },
         [&](bool ld) { tc2::gemm<true, true>(ld, tb + C_G2, XH2, DY16, 64, 16, 16, 8, 1u); });
    if (worker) bwd64(C_D0, mk_h2, DY64, false);  // background reads XH2 / DY16 only
    mark(11);
    step([&](bool ld) { tc2::gemm<false, true>(ld, tb + C_D0, DY64, W1, ROWS, 64, 64, 4, 0u); },
         [&](bool ld) { tc2::gemm<true, true>(ld, tb + C_G1, DY64, XH1, 64, 72, 64, 8, 1u); });
    // d(h1) goes into the XH2 buffer, dead since the GEMMs of the step before (ordered before this step's critical
    // GEMM) read it: no wait for this step's weight-gradient GEMM, which still reads DY64 and XH1
    if (worker) bwd64(C_D0, mk_h1, DH1, false);
    mark(12);
    // the next tile's rows have had three steps to arrive; the barrier of the step below publishes them
    if (gatherer) gather_wait();
    step_track(none, [&](bool ld) {
      tc2::gemm<true, true>(ld, tb + C_G0, DH1, XOBS, 64, 16, 8, 8, 1u);  // [dW^T | db]: obs c-group + the ones c-group
    });  // its commit also covers the GEMM on XH1 of the step before: waited for before XH1 / XH2 / XOBS are rewritten
    mark(13);
    swap_staging();
  }
  if (worker) {
    wait_crit();
    wait_bg();
  }

  mark(15);
  long long tf0 = 0, tf1 = 0, tf2 = 0, tf3 = 0;  // flush time stamps, kept in registers (a global read-modify-write per mark would stall the path it measures)
  if (prof_on) tf0 = clock64();
  // ------------------------------------------------------------------ flush the gradient partial
  // Every epilogue is done and every GEMM waited for: the activation buffers are free.  The dW^T accumulators (TMEM lane =
  // output unit, columns = inputs) are transposed through shared memory so that the partial is written with coalesced
  // stores: written straight from the TMEM fragments, every lane stored into a different 128-byte line -- ~13 k L1
  // store transactions per CTA, 13 k of the 15 k cycles this flush took (tools/tc128_prof.py), and the stores queued in
  // the LSU held up everything behind them.
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  float* part = a.partial + (size_t)blockIdx.x * a.partial_stride;
  const tc::LayerOffsets& lo = a.lo;
  const bool low = lane < 16;          // M = 64 accumulators: output / input unit mo at lanes 0-15 of the quadrant
  const int mo = q * 16 + (lane & 15);
  float* stg = reinterpret_cast<float*>(smem + OFF_XH1);  // XH1 | XH2: 66 KB, staging [64 outputs][in + 1] (odd pitch)
  constexpr int S_ENC1 = 0, S_V1 = 64 * 65, S_V0 = 2 * 64 * 65, S_A0 = S_V0 + 64 * 17, S_ENC0 = S_A0 + 64 * 17;
  static_assert((S_ENC0 + 64 * 9) * 4 <= 33 * CG, "flush staging exceeds the XH1 | XH2 region");
  // dW^T accumulators -> staging; the bias gradient (column bias_col) goes straight out (lanes = consecutive outputs)
  auto stage_T = [&](uint32_t col, int in_real, int in_pad, int bias_col, int layer, int soff, int pitch) {
    for (int c0 = 0; c0 < in_pad; c0 += 8) {
      float v[8];
      umma::tmem_ld8(tw + col + c0, v);
      umma::tmem_ld_wait();
      if (low) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c0 + j < in_real) stg[soff + mo * pitch + c0 + j] = v[j];
      }
    }
    float v[8];
    umma::tmem_ld8(tw + col + bias_col, v);
    umma::tmem_ld_wait();
    if (low) part[lo.b[layer] + mo] = v[0];
  };
  // dW accumulators: TMEM row = input unit, columns = outputs (their biases come from registers): lanes = consecutive
  // inputs of one output -> already coalesced
  auto flush_N = [&](uint32_t col, int in_real, int out_real, int out_pad, int layer) {
    for (int c0 = 0; c0 < out_pad; c0 += 8) {
      float v[8];
      umma::tmem_ld8(tw + col + c0, v);
      umma::tmem_ld_wait();
      if (low) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c0 + j < out_real && mo < in_real) part[lo.w[layer] + (c0 + j) * in_real + mo] = v[j];
      }
    }
  };
  if (worker && g == 0) { stage_T(C_G0, a.obs_dim, 8, 8, tc::L_ENC0, S_ENC0, 9); flush_N(C_G2, 64, 16, 16, tc::L_ENC2); }
  if (worker && g == 1) stage_T(C_G1, 64, 64, 64, tc::L_ENC1, S_ENC1, 65);
  if (worker && g == 2) { stage_T(C_GV1, 64, 64, 64, tc::L_V1, S_V1, 65); stage_T(C_GV0, 16, 16, 16, tc::L_V0, S_V0, 17); }
  if (worker && g == 3) { stage_T(C_GA0, 16, 16, 16, tc::L_A0, S_A0, 17); flush_N(C_GA1, 64, 16, 16, tc::L_A1); }

  // ------------------------------------------------------------------ register-held gradients and loss statistics
  float* sh = reinterpret_cast<float*>(smem + OFF_XB1);
  float* s_wv2 = sh;                 // [4 quadrants][64 columns]
  float* s_bv2 = s_wv2 + 256;        // [4 quadrants]
  float* s_e2 = s_bv2 + 4;           // [4 quadrants][16 columns]
  float* s_mw = s_e2 + 64;           // [16 warps][NO][16]
  float* s_mb = s_mw + 16 * LRX_MAX_NOUT * 16;  // [16 warps][NO]
  double* shd = reinterpret_cast<double*>(smem + OFF_DY64);  // [6][16 warps]
  if (worker) {
    if ((lane & 1) == 0) s_wv2[q * 64 + cg + (lane >> 1)] = gw_v2;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float x = warp_sum(bs_e2[j]);
      if (lane == 0) s_e2[q * 16 + 4 * g + j] = x;
    }
    if (g0) {
      const float x = warp_sum(gb_v2);
      if (lane == 0) s_bv2[q] = x;
    }
    // mapping layer: reduce over the warp's eight row slots (lanes with equal lane % 4), one slot per warp
#pragma unroll
    for (int o2 = 0; o2 < NO; ++o2) {
      if (o2 < n_out) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float x = gm_w[o2][c];
          x += __shfl_xor_sync(0xffffffffu, x, 4);
          x += __shfl_xor_sync(0xffffffffu, x, 8);
          x += __shfl_xor_sync(0xffffffffu, x, 16);
          if (lane < 4) s_mw[(warp * LRX_MAX_NOUT + o2) * 16 + 4 * lane + c] = x;
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
  if (prof_on) tf1 = clock64();
  __syncthreads();
  if (prof_on) tf2 = clock64();
  // ------------------------------------------------------------------ staging -> partial, coalesced; fixed-order sums
  {
    // 64 x 64 layers: shifts instead of a division by a run-time width; consecutive threads = consecutive addresses
    for (int e = tid; e < 2 * 4096; e += THREADS) {
      const int l2 = e >> 12, r = e & 4095, o = r >> 6, i = r & 63;
      part[lo.w[l2 ? tc::L_V1 : tc::L_ENC1] + r] = stg[(l2 ? S_V1 : S_ENC1) + o * 65 + i];
    }
    for (int e = tid; e < 2 * 1024; e += THREADS) {
      const int l2 = e >> 10, r = e & 1023, o = r >> 4, i = r & 15;
      part[lo.w[l2 ? tc::L_A0 : tc::L_V0] + r] = stg[(l2 ? S_A0 : S_V0) + o * 17 + i];
    }
    for (int e = tid; e < 64 * a.obs_dim; e += THREADS) {
      const int o = e / a.obs_dim, i = e - o * a.obs_dim;
      part[lo.w[tc::L_ENC0] + e] = stg[S_ENC0 + o * 9 + i];
    }
    auto sum4 = [](const float* p, int stride) { return ((p[0] + p[stride]) + p[2 * stride]) + p[3 * stride]; };
    auto sum16 = [](const float* p, int stride) {
      float t = 0.f;
      for (int w = 0; w < 16; ++w) t += p[w * stride];  // warp order: fixed
      return t;
    };
    if (tid < 64) part[lo.w[tc::L_V2] + tid] = sum4(s_wv2 + tid, 64);
    if (tid == 64) part[lo.b[tc::L_V2]] = sum4(s_bv2, 1);
    if (tid >= 96 && tid < 112) part[lo.b[tc::L_ENC2] + tid - 96] = sum4(s_e2 + tid - 96, 16);
    if (tid >= 128 && tid < 128 + 16 * LRX_MAX_NOUT) {
      const int o2 = (tid - 128) >> 4, i2 = (tid - 128) & 15;
      if (o2 < n_out) part[lo.w[tc::L_MAP] + o2 * 16 + i2] = sum16(s_mw + o2 * 16 + i2, LRX_MAX_NOUT * 16);
    }
    if (tid >= 256 && tid < 256 + LRX_MAX_NOUT && tid - 256 < n_out)
      part[lo.b[tc::L_MAP] + tid - 256] = sum16(s_mb + tid - 256, LRX_MAX_NOUT);
    if (tid >= 288 && tid < 304) {  // db_a1 = W_map^T db_map (the layers are linear in between)
      const int i2 = tid - 288;
      float sacc = 0.f;
      for (int o2 = 0; o2 < n_out; ++o2) sacc = fmaf(sum16(s_mb + o2, LRX_MAX_NOUT), fp[tc::F_WMAP + o2 * 16 + i2], sacc);
      part[lo.b[tc::L_A1] + i2] = sacc;
    }
    if (tid >= 320 && tid < 326) {
      const int k = tid - 320;
      double s2 = 0;
      for (int w = 0; w < 16; ++w) s2 += shd[k * 16 + w];
      a.loss_partial[(size_t)blockIdx.x * 8 + k] = s2;
    }
  }
  if (prof_on) {
    tf3 = clock64();
    const int o = tid == 0 ? 64 : 96;
    a.prof[o + 0] += tf1 - tf0; a.prof[o + 1] += tf2 - tf1; a.prof[o + 2] += tf3 - tf2;
  }
  mark(16);
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

static long long* g_tc128_prof = nullptr;  // test hook, include/lerax_b200_debug.h
extern "C" void lrx_debug_set_tc128_prof(long long* device_buffer_64) { g_tc128_prof = device_buffer_64; }
long long* lrx_tc128_prof_ptr() { return g_tc128_prof; }

// Same workspace layout as the first-generation kernel (lrx_fused_workspace_bytes covers both).
int lrx_ppo_grad_tc128(const LrxPolicyDesc* pol, const float* params, const LrxBatchView* buf, const int32_t* idx,
                       int64_t B, int64_t B_global, const double* adv_sums, const LrxPpoHyper* h, float* gradbuf,
                       void* workspace, size_t workspace_bytes, cudaStream_t stream, bool* handled,
                       const LrxApplyArgs* apply) {
  *handled = false;
  if (!tc::is_default_policy(*pol) || lrx_fused_enabled() < 2) return LRX_OK;
  FusedWorkspace w = carve_fused(*pol, workspace);
  LRX_REQUIRE(workspace_bytes >= w.total, LRX_ERR_INVALID_ARG, "lrx_ppo_grad: workspace too small for the fused path");
  Tc128Args a{};
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
  a.prof = g_tc128_prof;
  const int sms = lrx_sm_count();
  const int grid = a.n_tiles < sms ? a.n_tiles : sms;
  static bool set0[64], set1[64], set2[64], set3[64];
  LRX_CUDA_CHECK(lrx_set_max_smem(ppo_grad_tc128_kernel<0>, (int)SMEM_BYTES, set0));
  LRX_CUDA_CHECK(lrx_set_max_smem(ppo_grad_tc128_kernel<1>, (int)SMEM_BYTES, set1));
  LRX_CUDA_CHECK(lrx_set_max_smem(ppo_grad_tc128_kernel<2>, (int)SMEM_BYTES, set2));
  LRX_CUDA_CHECK(lrx_set_max_smem(ppo_grad_tc128_kernel<3>, (int)SMEM_BYTES, set3));
  switch (pol->n_out) {
    case 1: ppo_grad_tc128_kernel<1><<<grid, THREADS, SMEM_BYTES, stream>>>(a); break;
    case 2: ppo_grad_tc128_kernel<2><<<grid, THREADS, SMEM_BYTES, stream>>>(a); break;
    case 3: ppo_grad_tc128_kernel<3><<<grid, THREADS, SMEM_BYTES, stream>>>(a); break;
    default: ppo_grad_tc128_kernel<0><<<grid, THREADS, SMEM_BYTES, stream>>>(a); break;
  }
  LRX_LAUNCH_CHECK();
  const int rc = lrx_launch_reduce_partials(w.partial, grid, (long long)w.P_pad, pol->n_params, w.loss_partial, grid,
                                            a.lo.log_std, h->entropy_loss_coefficient, (double)B_global, gradbuf, apply, stream);
  if (rc) return rc;
  LRX_LAUNCH_CHECK();
  *handled = true;
  return LRX_OK;
}
