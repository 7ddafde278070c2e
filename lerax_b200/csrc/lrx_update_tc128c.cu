// Fused PPO minibatch gradient for the default MLPActorCriticPolicy widths, third revision of the 128-row tensor-core
// kernel (lrx_update_tc128.cu): COMPOSITE WEIGHTS across the policy's linear-linear seams.
//
// Reference: algorithm/ppo.py:396-469 (ppo_loss + filter_value_and_grad) over the gathered minibatch of
// buffer/base_buffer.py:90-103; policy/actor_critic/mlp.py:64-178 (encoder d->64->64->16 WITHOUT an activation on its
// 16-wide output, heads 16->64->..), policy/actor.py:248-261 (action_dist.mapping, a Linear right after the action
// head's linear output layer).
//
// Why.  lrx_update_tc128.cu is bound by its serial chain of twelve MMA -> commit -> tcgen05.ld -> epilogue -> barrier
// steps per 128-row tile (tools/tc128_prof.py: 25.6 k cycles per tile, the tensor pipe ~35 % occupied), and the longest
// step was the policy loss spread over 4 lanes per row (4.3 k cycles: sixteen shuffles, a divergent select and the
// 16 -> n_out mapping on the FMA pipe in front of the loss itself).  Two of the network's seams have NO nonlinearity:
//     features f = W2 h2 + b2   ->  head layer 0:  relu(Wv0 f + bv0) = relu((Wv0 W2) h2 + (Wv0 b2 + bv0))
//     a1 = Wa1 xa + ba1         ->  mapping:       logits = Wmap a1 + bmap = (Wmap Wa1) xa + (Wmap ba1 + bmap)
// so the kernel multiplies by the composites Wcv = Wv0 W2, Wca = Wa0 W2 (64 x 64) and Wcm = Wmap Wa1 (n_out x 64), built
// once per launch in fp32, and never forms f, d(f) or a1 per row:
//   * forward: enc2 disappears from the chain (heads read h2 directly); the action head's output IS the logits, so the
//     policy loss is one thread per row on the four row-owner warps, no shuffles, no mapping arithmetic;
//   * backward: d(h2) = dv0 Wcv + da0 Wca accumulates in ONE TMEM tile (the value head's half is a background GEMM under the
//     policy loss), replacing the d(features) step and the enc2 backward step;
//   * weight gradients are linear in the TMEM accumulators Gv = dv0^T [h2 | 1], Ga = da0^T [h2 | 1], Gm = xa^T dlogits:
//         dWv0 = Gv_h W2^T + gv_1 b2^T,  dbv0 = gv_1            (same for a0 with Ga)
//         dW2  = Wv0^T Gv_h + Wa0^T Ga_h, db2 = Wv0^T gv_1 + Wa0^T ga_1
//         dWa1 = Wmap^T Gm^T,  dba1 = Wmap^T gm_1,  dWmap = Gm^T Wa1^T + gm_1 ba1^T,  dbmap = gm_1
//     evaluated once per CTA at the flush in fp32 (280 k MAC), so the per-CTA partial keeps its layout.
// Ten barrier steps per tile instead of twelve.  Same arithmetic contract as lrx_update_tc128.cu (two bf16 pieces, three
// products, 1e-4 on gradients); the composites add one fp32 rounding of a 16-term dot product per weight.
//
// The minibatch rows arrive as 64-byte packed records (lrx_ppo_pack_rows) through four 16-byte cp.async per row into a
// single staging buffer, tracked by an mbarrier (cp.async.mbarrier.arrive); plane-wise buffers are gathered field by field.
// (cp.async.bulk per row was measured and dropped: ~47 cycles per 32-byte copy for the issuing warp, 3 k cycles per tile.)
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
constexpr uint32_t OFF_WCV = OFF_W1 + 16384;         // [64][64]  Wv0 W2
constexpr uint32_t OFF_WCA = OFF_WCV + 16384;        // [64][64]  Wa0 W2
constexpr uint32_t OFF_WV1 = OFF_WCA + 16384;        // [64][64]
constexpr uint32_t OFF_WCM = OFF_WV1 + 16384;        // [16][64]  Wmap Wa1 (rows >= n_out zero): hi 2048 | lo 2048
constexpr uint32_t OFF_XOBS = OFF_WCM + 4096;        // [128][8 obs]: hi 1 c-group, lo 1; its ones column is XH1's
constexpr uint32_t OFF_XH1 = OFF_XOBS + 2 * CG;      // [128][64 | ones]: hi 9 c-groups, lo 8
constexpr uint32_t OFF_XH2 = OFF_XH1 + 17 * CG;      // [128][64 | ones]
constexpr uint32_t OFF_XB1 = OFF_XH2 + 17 * CG;      // [128][64 | ones]: head layer-0 activations (value, then action); d(h2) at the end
constexpr uint32_t OFF_DY64 = OFF_XB1 + 17 * CG;     // [128][64]: current upstream gradient
constexpr uint32_t OFF_DY16 = OFF_DY64 + 16 * CG;    // [128][16]: d(logits), columns >= 8 stay zero
constexpr uint32_t OFF_VPART = OFF_DY16;             // [4 column groups][128 rows] value dot partials: dead before DY16's c-group 0 is written
// fp32 staging of the gathered rows (single-buffered: consumed by the end of the policy loss, refilled right after)
constexpr uint32_t OFF_RECA = OFF_DY16 + 4 * CG;     // [128][8]  observation (record floats 0..7)
constexpr uint32_t OFF_RECB = OFF_RECA + ROWS * 32;  // [128][8]  action bits, old log-prob, old value, return | advantage, 0, 0, 0
// fp32 parameter block
enum : int {
  FB0 = 0, FB1 = 64, FBCV = 128, FBV1 = 192, FBCA = 256, FBV2 = 320, FBCM = 324, FWV2 = 332, FW0 = 396,
  F_COUNT = 908, F_RESERVE = 912
};
constexpr uint32_t OFF_F32 = OFF_RECB + ROWS * 32;
constexpr uint32_t OFF_MISC = OFF_F32 + F_RESERVE * 4;   // 4 mbarriers, TMEM slot
constexpr uint32_t OFF_IDX = OFF_MISC + 64;              // minibatch indices of the tile after the one being gathered
constexpr uint32_t SMEM_BYTES = OFF_IDX + ROWS * 4;
static_assert(SMEM_BYTES <= 227 * 1024, "tc128c update kernel exceeds shared memory");

// ---- TMEM columns (fp32).  M = 128 accumulators use all 128 lanes; the M = 64 weight-gradient accumulators use
// lanes 0-15 of every 32-lane quadrant (output unit j at lane 32 * (j / 16) + j % 16).
constexpr uint32_t C_D0 = 0;       // 64: layer outputs
constexpr uint32_t C_D1 = 64;      // 64: action head layer-0 pre-activations (issued with the value head's) ...
constexpr uint32_t C_DH2 = 64;     // ... later d(h2), summed over both heads
constexpr uint32_t C_G0 = 128;     // enc0  dW^T [64 out][8 in | bias | 0 x 7]
constexpr uint32_t C_G1 = 144;     // enc1  dW^T [64][64 | bias, 0 x 7]
constexpr uint32_t C_GV1 = 216;    // v1    dW^T [64][64 | bias, 0 x 7]
constexpr uint32_t C_GV = 288;     // Gv = dv0^T [h2 | 1]
constexpr uint32_t C_GA = 360;     // Ga = da0^T [h2 | 1]
constexpr uint32_t C_GM = 432;     // Gm = xa^T dlogits [64 in][16]
constexpr uint32_t TMEM_COLS = 512;

struct Tc128cArgs {
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

// arrive on `bar` once every cp.async this thread has issued so far has landed (counted in the barrier's init count)
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}

// NOUT: number of action-head outputs when known at compile time (1 = scalar Box, 2 = CartPole, 3 = Acrobot /
// MountainCar); 0 = read it from the arguments (up to LRX_MAX_NOUT).
template <int NOUT>
__global__ void __launch_bounds__(THREADS, 1) ppo_grad_tc128c_kernel(const __grid_constant__ Tc128cArgs a) {
  constexpr int NO = NOUT ? NOUT : LRX_MAX_NOUT;
  const int n_out = NOUT ? NOUT : a.n_out;
  extern __shared__ __align__(1024) uint8_t smem[];
  // Programmatic dependent launch: this grid may be scheduled while the previous kernel of the stream (the reduce / clip /
  // Adam kernel that writes the parameters) is still draining; nothing of its output is touched before the wait, and the
  // next kernel (the reduction of this grid's partials) may be set up as early as it likes -- it waits the same way.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
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
  uint64_t* barR = reinterpret_cast<uint64_t*>(smem + OFF_MISC + 32);  // the next tile's rows have landed
  int* idx_s = reinterpret_cast<int*>(smem + OFF_IDX);
  float* recA = reinterpret_cast<float*>(smem + OFF_RECA);
  float* recB = reinterpret_cast<float*>(smem + OFF_RECB);

  const tc2::Mat W1{smem + OFF_W1, 64, 8192}, WCV{smem + OFF_WCV, 64, 8192}, WCA{smem + OFF_WCA, 64, 8192};
  const tc2::Mat WV1{smem + OFF_WV1, 64, 8192}, WCM{smem + OFF_WCM, 16, 2048};
  // [obs | 1]: the second c-group of the hi piece is the ones c-group of XH1 (10 c-groups further on)
  const tc2::Mat XOBS{smem + OFF_XOBS, ROWS, CG, (OFF_XH1 + 8 * CG) - OFF_XOBS};
  const tc2::Mat XH1{smem + OFF_XH1, ROWS, 9 * CG}, XH2{smem + OFF_XH2, ROWS, 9 * CG};
  const tc2::Mat XB1{smem + OFF_XB1, ROWS, 9 * CG};
  const tc2::Mat DY64{smem + OFF_DY64, ROWS, 8 * CG}, DY16{smem + OFF_DY16, ROWS, 2 * CG};

  // ------------------------------------------------------------------ one-time staging
  {
    // W1 and Wv1: 1024 chunks of 8 consecutive inputs of one output unit, loads issued before any conversion
    constexpr int NCH = 1024, PER = (NCH + THREADS - 1) / THREADS;
    float v[PER][8];
    int oo[PER], cc[PER], mi[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int e = tid + k * THREADS;
      mi[k] = -1;
      if (e < NCH) {
        mi[k] = e >> 9;
        const int el = e & 511;
        oo[k] = el & 63; cc[k] = el >> 6;
        const float* src = a.params + a.lo.w[mi[k] ? tc::L_V1 : tc::L_ENC1] + (size_t)oo[k] * 64 + cc[k] * 8;
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
          const float4 x = __ldg(reinterpret_cast<const float4*>(src)), y = __ldg(reinterpret_cast<const float4*>(src) + 1);
          v[k][0] = x.x; v[k][1] = x.y; v[k][2] = x.z; v[k][3] = x.w; v[k][4] = y.x; v[k][5] = y.y; v[k][6] = y.z; v[k][7] = y.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[k][j] = __ldg(src + j);
        }
      }
    }
    // fp32 sources of the composites, in the (still unused) DY64 region:
    //   W2 [16][64] | Wv0 [64][16] | Wa0 [64][16] | Wa1 [16][64] | Wmap [8][16] (rows >= n_out zero) | b2 [16] | ba1 [16]
    float* cs = reinterpret_cast<float*>(smem + OFF_DY64);
    constexpr int CS_W2 = 0, CS_WV0 = 1024, CS_WA0 = 2048, CS_WA1 = 3072, CS_WMAP = 4096, CS_B2 = 4224, CS_BA1 = 4240, CS_N = 4256;
    // Source address by selection, every load of the thread issued before the first store: as a loop of load -> store
    // with a branch per range (what the compiler makes of the obvious form) these were ten dependent L2 round trips per
    // launch.  nullptr = the entry is a constant zero (skip = false) or is not written here (skip = true).
    constexpr int CS_PER = (CS_N + THREADS - 1) / THREADS, FP_PER = (F_COUNT + THREADS - 1) / THREADS;
    float xc[CS_PER], xf[FP_PER];
    bool skip_f[FP_PER];
#pragma unroll
    for (int k = 0; k < CS_PER; ++k) {
      const int i = tid + k * THREADS;
      const float* src = nullptr;
      if (i < CS_WV0) src = a.params + a.lo.w[tc::L_ENC2] + i;
      else if (i < CS_WA0) src = a.params + a.lo.w[tc::L_V0] + (i - CS_WV0);
      else if (i < CS_WA1) src = a.params + a.lo.w[tc::L_A0] + (i - CS_WA0);
      else if (i < CS_WMAP) src = a.params + a.lo.w[tc::L_A1] + (i - CS_WA1);
      else if (i < CS_B2) src = (i - CS_WMAP) < n_out * 16 ? a.params + a.lo.w[tc::L_MAP] + (i - CS_WMAP) : nullptr;
      else if (i < CS_BA1) src = a.params + a.lo.b[tc::L_ENC2] + (i - CS_B2);
      else if (i < CS_N) src = a.params + a.lo.b[tc::L_A1] + (i - CS_BA1);
      xc[k] = src ? __ldg(src) : 0.f;
    }
    // fp32 parameter block, except the composite biases
#pragma unroll
    for (int k = 0; k < FP_PER; ++k) {
      const int i = tid + k * THREADS;
      const float* src = nullptr;
      bool skip = i >= F_COUNT;
      if (i < FB1) src = a.params + a.lo.b[tc::L_ENC0] + i;
      else if (i < FBCV) src = a.params + a.lo.b[tc::L_ENC1] + (i - FB1);
      else if (i < FBV1) skip = true;
      else if (i < FBCA) src = a.params + a.lo.b[tc::L_V1] + (i - FBV1);
      else if (i < FBV2) skip = true;
      else if (i < FBCM) src = (i == FBV2) ? a.params + a.lo.b[tc::L_V2] : nullptr;
      else if (i < FWV2) skip = true;
      else if (i < FW0) src = a.params + a.lo.w[tc::L_V2] + (i - FWV2);
      else if (i < F_COUNT) {
        const int o = (i - FW0) >> 3, kk = (i - FW0) & 7;
        src = kk < a.obs_dim ? a.params + a.lo.w[tc::L_ENC0] + o * a.obs_dim + kk : nullptr;
      }
      xf[k] = (src && !skip) ? __ldg(src) : 0.f;
      skip_f[k] = skip;
    }
#pragma unroll
    for (int k = 0; k < CS_PER; ++k)
      if (tid + k * THREADS < CS_N) cs[tid + k * THREADS] = xc[k];
#pragma unroll
    for (int k = 0; k < FP_PER; ++k)
      if (!skip_f[k]) fp[tid + k * THREADS] = xf[k];
#pragma unroll
    for (int k = 0; k < PER; ++k)
      if (mi[k] >= 0) tc2::store8(mi[k] ? WV1 : W1, cc[k], oo[k], v[k]);
    __syncthreads();
    // composites: chunk e = 8 consecutive inputs i of one output o:  Wc[o][i] = sum_k L[o][k] R[k][i]  (k = 0..15 in order)
    for (int e = tid; e < 512 + 512 + 128; e += THREADS) {
      const int mat = e < 512 ? 0 : (e < 1024 ? 1 : 2);
      const int el = e - (mat == 2 ? 1024 : mat * 512);
      const int o = mat == 2 ? (el & 15) : (el & 63), ch = mat == 2 ? (el >> 4) : (el >> 6);
      const float* Lrow = cs + (mat == 0 ? CS_WV0 : (mat == 1 ? CS_WA0 : CS_WMAP)) + o * 16;  // rows 8..15 of Wmap: see below
      const float* R = cs + (mat == 2 ? CS_WA1 : CS_W2) + ch * 8;
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (mat != 2 || o < LRX_MAX_NOUT) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float l = Lrow[k];
          const float4 r0 = *reinterpret_cast<const float4*>(R + k * 64), r1 = *reinterpret_cast<const float4*>(R + k * 64 + 4);
          acc[0] = fmaf(l, r0.x, acc[0]); acc[1] = fmaf(l, r0.y, acc[1]); acc[2] = fmaf(l, r0.z, acc[2]); acc[3] = fmaf(l, r0.w, acc[3]);
          acc[4] = fmaf(l, r1.x, acc[4]); acc[5] = fmaf(l, r1.y, acc[5]); acc[6] = fmaf(l, r1.z, acc[6]); acc[7] = fmaf(l, r1.w, acc[7]);
        }
      }
      tc2::store8(mat == 0 ? WCV : (mat == 1 ? WCA : WCM), ch, o, acc);
    }
    // composite biases
    if (tid < 64 + 64 + LRX_MAX_NOUT) {
      const int mat = tid < 64 ? 0 : (tid < 128 ? 1 : 2), o = tid - (mat == 2 ? 128 : mat * 64);
      const float* Lrow = cs + (mat == 0 ? CS_WV0 : (mat == 1 ? CS_WA0 : CS_WMAP)) + o * 16;
      const float* bb = cs + (mat == 2 ? CS_BA1 : CS_B2);
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) acc = fmaf(Lrow[k], bb[k], acc);
      if (mat == 2) acc = o < n_out ? acc + __ldg(a.params + a.lo.b[tc::L_MAP] + o) : 0.f;
      else acc += __ldg(a.params + a.lo.b[mat ? tc::L_A0 : tc::L_V0] + o);
      fp[(mat == 0 ? FBCV : (mat == 1 ? FBCA : FBCM)) + o] = acc;
    }
    __syncthreads();  // the composite sources in DY64 are dead from here on
  }
  if (tid < ROWS) {  // the constant ones columns (bf16 1.0 = 0x3F80 in element 0 of the extra c-group); d(logits) columns 8..15 = 0
    const uint4 one = make_uint4(0x00003F80u, 0u, 0u, 0u), zero = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(XH1.ptr + 8 * CG + tid * 16) = one;
    *reinterpret_cast<uint4*>(XH2.ptr + 8 * CG + tid * 16) = one;
    *reinterpret_cast<uint4*>(XB1.ptr + 8 * CG + tid * 16) = one;
    *reinterpret_cast<uint4*>(DY16.ptr + CG + tid * 16) = zero;
    *reinterpret_cast<uint4*>(DY16.ptr + DY16.lo_off + CG + tid * 16) = zero;
  }
  if (tid == 0) {
    umma::mbar_init(barA, 1);
    umma::mbar_init(barB, 1);
    umma::mbar_init(barL, WORKERS / 32);
    umma::mbar_init(barR, ROWS);
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
  uint32_t phA = 0, phB = 0, phL = 0, phR = 0;
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

  // ---- row gather (buffer/base_buffer.py:90-103), one tile ahead, by the warps of column group 2 (one lane per row).
  // LDGSTS copies (zero-size = zero fill), cp.async.mbarrier.arrive on barR.  Rows past the minibatch are zero filled (finite operands for the
  // weight-gradient GEMMs; their loss gradients are zero).
  auto cp_async4 = [](void* dst, const void* src, bool ok) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(umma::smem_u32(dst)), "l"(src), "r"(ok ? 4 : 0) : "memory");
  };
  auto cp_async16 = [](void* dst, const void* src, bool ok) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(umma::smem_u32(dst)), "l"(src), "r"(ok ? 16 : 0) : "memory");
  };
  auto tile_row_ok = [&](int t, int r) { return t < a.n_tiles && (long long)t * ROWS + r < a.B; };
  auto gather_tile = [&](int t_next, int t_after, bool first) {
    const bool ok = tile_row_ok(t_next, m);
    int i = 0;
    if (first) {
      if (ok) i = __ldg(a.idx + (long long)t_next * ROWS + m);
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      i = idx_s[m];
    }
    float* ra = recA + m * 8;
    float* rb = recB + m * 8;
    if (a.buf.packed != nullptr) {
      // row-major record (lrx_ppo_pack_rows): one 64-byte line per row instead of one sector per plane
      const float* rec = a.buf.packed + (ok ? (size_t)i * LRX_PACKED_FLOATS : 0);
      cp_async16(ra, rec, ok);
      cp_async16(ra + 4, rec + 4, ok);
      cp_async16(rb, rec + 8, ok);
      cp_async16(rb + 4, rec + 12, ok);
      cp_async_arrive_noinc(barR);
    } else {
      const int n = i / a.buf.T, t = i - n * a.buf.T;
      const size_t src = ok ? (size_t)t * a.buf.N + n : 0;
      cp_async4(rb + 0, (const int32_t*)a.buf.act + src, ok);
      cp_async4(rb + 1, a.buf.logp + src, ok);
      cp_async4(rb + 2, a.buf.val + src, ok);
      cp_async4(rb + 3, a.buf.ret + src, ok);
      cp_async4(rb + 4, a.buf.adv + src, ok);  // raw; normalised where it is used
      const float* op = a.buf.obs + src * a.obs_dim;
#pragma unroll
      for (int k = 0; k < 8; ++k) cp_async4(ra + k, op + (k < a.obs_dim ? k : 0), ok && k < a.obs_dim);
      cp_async_arrive_noinc(barR);
    }
    const bool ok2 = tile_row_ok(t_after, m);  // minibatch index of the tile after: read only by this lane, next time
    cp_async4(idx_s + m, a.idx + (ok2 ? (long long)t_after * ROWS + m : 0), ok2);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const bool gatherer = worker && g == 2;
  if (gatherer) gather_tile((int)blockIdx.x, (int)blockIdx.x + (int)gridDim.x, true);

  // gradients kept in registers across the tiles of the CTA
  float gw_v2 = 0.f, gb_v2 = 0.f;  // value output layer (64 -> 1): dW of column cg + (lane >> 1) % 16 over this warp's rows; db
  float gm_b[NO];                  // column sums of d(logits) over this thread's rows (row-owner warps)
#pragma unroll
  for (int o2 = 0; o2 < NO; ++o2) gm_b[o2] = 0.f;

  mark(0);
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    mark(14);
    const bool valid_m = (long long)tile * ROWS + m < a.B;  // this thread's row is a real row of the minibatch
    mark(0);

    // ---------------------------------------------------------------- encoder layer 0 (d <= 8 inputs) on the FMA pipe
    uint32_t mk_h1 = 0, mk_h2 = 0, mk_b1 = 0, mk_b2 = 0;
    if (worker) {
      umma::mbar_wait(barR, phR);  // this tile's rows
      phR ^= 1u;
      float x[8];
      {
        const float4 lo4 = *reinterpret_cast<const float4*>(recA + m * 8), hi4 = *reinterpret_cast<const float4*>(recA + m * 8 + 4);
        x[0] = lo4.x; x[1] = lo4.y; x[2] = lo4.z; x[3] = lo4.w; x[4] = hi4.x; x[5] = hi4.y; x[6] = hi4.z; x[7] = hi4.w;
      }
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 wa = *reinterpret_cast<const float4*>(fp + FW0 + (cg + j) * 8);  // warp-uniform: broadcast
        float acc = fmaf(wa.w, x[3], fmaf(wa.z, x[2], fmaf(wa.y, x[1], wa.x * x[0])));
        if (a.obs_dim > 4) {
          const float4 wb = *reinterpret_cast<const float4*>(fp + FW0 + (cg + j) * 8 + 4);
          acc = fmaf(wb.w, x[7], fmaf(wb.z, x[6], fmaf(wb.y, x[5], fmaf(wb.x, x[4], acc))));
        }
        v[j] = acc;
      }
      add_bias_relu(v, fp + FB0, mk_h1);
      store_row16(XH1, v, true);  // waits for the previous tile's last GEMMs (they read XH1 / XB1 / DY64 / XOBS)
      if (g0) {
        float z8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) z8[k] = k < a.obs_dim ? x[k] : 0.f;
        tc2::store8(XOBS, 0, m, z8);
      }
    }
    mark(1);
    step([&](bool ld) { tc2::gemm<false, false>(ld, tb + C_D0, XH1, W1, ROWS, 64, 64, 4, 0u); }, none);
    if (worker) fwd64(C_D0, fp + FB1, XH2, mk_h2);
    mark(2);
    // ---------------------------------------------------------------- both heads' layer 0 straight from h2 (composites)
    step([&](bool ld) { tc2::gemm<false, false>(ld, tb + C_D0, XH2, WCV, ROWS, 64, 64, 4, 0u); },
         [&](bool ld) { tc2::gemm<false, false>(ld, tb + C_D1, XH2, WCA, ROWS, 64, 64, 4, 0u); });  // action head, consumed later
    if (worker) fwd64(C_D0, fp + FBCV, XB1, mk_b1);
    mark(3);
    step([&](bool ld) { tc2::gemm<false, false>(ld, tb + C_D0, XB1, WV1, ROWS, 64, 64, 4, 0u); }, none);
    float xb2[16];  // value layer-1 activations of this thread's row and column group
    float4 rec4 = make_float4(0.f, 0.f, 0.f, 0.f);  // action bits, old log-prob, old value, return of this thread's row
    float adv_r = 0.f;                              // its raw advantage
    if (worker) {
      wait_crit();
      umma::tmem_ld16(tw + C_D0 + cg, xb2);
      umma::tmem_ld_wait();
      add_bias_relu(xb2, fp + FBV1, mk_b2);
      const float4* wp = reinterpret_cast<const float4*>(fp + FWV2 + cg);
      float pr = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 w = wp[i];
        pr = fmaf(w.w, xb2[4 * i + 3], fmaf(w.z, xb2[4 * i + 2], fmaf(w.y, xb2[4 * i + 1], fmaf(w.x, xb2[4 * i], pr))));
      }
      vpart[g * ROWS + m] = pr;
      rec4 = *reinterpret_cast<const float4*>(recB + m * 8);
      adv_r = recB[m * 8 + 4];
    }
    __syncthreads();

    mark(4);
    // ---------------------------------------------------------------- value loss (ppo.py:435-451) and its backward
    if (worker) {
      const float value = (((fp[FBV2] + vpart[m]) + vpart[ROWS + m]) + vpart[2 * ROWS + m]) + vpart[3 * ROWS + m];
      const float ret_r = rec4.w, vold_r = rec4.z;
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
      const float4* wp = reinterpret_cast<const float4*>(fp + FWV2 + cg);
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
    mark(5);
    step_track([&](bool ld) { tc2::gemm<false, true>(ld, tb + C_D0, DY64, WV1, ROWS, 64, 64, 4, 0u); },      // dX = dY W
               [&](bool ld) { tc2::gemm<true, true>(ld, tb + C_GV1, DY64, XB1, 64, 72, 64, 8, 1u); });        // [dW^T | db] += dY^T [X | 1]
    // The staged rows of this tile were consumed before the barrier above (observations at the top, the per-row scalars into
    // registers in the value step): fetch the next tile's now, at the start of the longest stretch between two
    // fence.proxy.async (which drain the CTA's outstanding LDGSTS: a copy still in flight at the next fence stalls it).
    if (gatherer) gather_tile(tile + (int)gridDim.x, tile + 2 * (int)gridDim.x, false);
    if (worker) {
      bwd64(C_D0, mk_b1, DY64, true);  // dv0; the background GEMM reads DY64 and XB1
      // -------------------------------------------------------------- action head forward
      fwd64(C_D1, fp + FBCA, XB1, mk_b1);
    }
    mark(6);
    // logits = xa Wcm^T.  Background, issued once the logits have been loaded from TMEM, under the policy loss: the value
    // head's share of d(h2) and Gv
    step_late([&](bool ld) { tc2::gemm<false, false>(ld, tb + C_D0, XB1, WCM, ROWS, 16, 16, 4, 0u); },
              [&](bool ld) {
                tc2::gemm<false, true>(ld, tb + C_DH2, DY64, WCV, ROWS, 64, 64, 4, 0u);
                tc2::gemm<true, true>(ld, tb + C_GV, DY64, XH2, 64, 72, 64, 8, 1u);
              });
    if (worker) {
      // ------------------------------------------------------------ policy loss (ppo.py:405-434,452-467) and its backward:
      // one thread per row on the four row-owner warps (one per scheduler)
      wait_crit();
      float lg[8];
      if (g0) {
        umma::tmem_ld8(tw + C_D0, lg);
        umma::tmem_ld_wait();
      }
      signal_loaded();
      if (g0) {
        float head[NO], dhead[NO];
#pragma unroll
        for (int o2 = 0; o2 < NO; ++o2) {
          head[o2] = o2 < n_out ? lg[o2] + fp[FBCM + o2] : 0.f;
          dhead[o2] = 0.f;
        }
        if (valid_m) {
          const int act_raw = __float_as_int(rec4.x);  // fp32 (Box) or int32 (Discrete) bits
          const float logp_old = rec4.y, A = adv_r;
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
          if (bad_p) s_bad += 1.0f;
          s_kl += ratio - log_ratio;
          s_pol += spol;
          s_ent += entropy;
          s_dls += sdls;
#pragma unroll
          for (int o2 = 0; o2 < NO; ++o2) gm_b[o2] += dhead[o2];
        }
        float d8[8];
#pragma unroll
        for (int o2 = 0; o2 < 8; ++o2) d8[o2] = 0.f;
#pragma unroll
        for (int o2 = 0; o2 < NO; ++o2) d8[o2] = dhead[o2];
        tc2::store8(DY16, 0, m, d8);  // columns 8..15 were zeroed once
      }
    }
    mark(7);
    step([&](bool ld) { tc2::gemm<false, true>(ld, tb + C_D0, DY16, WCM, ROWS, 64, 64, 1, 0u); },        // d(xa) = dlogits Wcm (K = 16)
         [&](bool ld) { tc2::gemm<true, true>(ld, tb + C_GM, XB1, DY16, 64, 16, 16, 8, 1u); });           // Gm[in][out] += xa^T dlogits
    if (worker) bwd64(C_D0, mk_b1, DY64, false);  // da0; DY64's readers (the step before) are ordered before this step's GEMM
    mark(8);
    // ---------------------------------------------------------------- encoder backward
    step([&](bool ld) { tc2::gemm<false, true>(ld, tb + C_DH2, DY64, WCA, ROWS, 64, 64, 4, 1u); },       // d(h2) += action head's
         [&](bool ld) { tc2::gemm<true, true>(ld, tb + C_GA, DY64, XH2, 64, 72, 64, 8, 1u); });
    if (worker) bwd64(C_DH2, mk_h2, XB1, false);  // d(h2 pre-activation) -> XB1 (its last reader was the step before)
    mark(9);
    step([&](bool ld) { tc2::gemm<false, true>(ld, tb + C_D0, XB1, W1, ROWS, 64, 64, 4, 0u); },
         [&](bool ld) { tc2::gemm<true, true>(ld, tb + C_G1, XB1, XH1, 64, 72, 64, 8, 1u); });
    if (worker) bwd64(C_D0, mk_h1, DY64, false);  // d(h1 pre-activation) -> DY64 (read by Ga's GEMM, ordered before this step's)
    mark(10);
    step_track(none, [&](bool ld) {
      tc2::gemm<true, true>(ld, tb + C_G0, DY64, XOBS, 64, 16, 8, 8, 1u);  // [dW^T | db]: obs c-group + the shared ones c-group
    });  // its commit also covers the GEMM on XH1 / XB1 of the step before: waited for before XH1 is rewritten
    mark(11);
  }
  if (worker) {
    wait_crit();
    wait_bg();
  }
  if (gatherer) {  // the tail gather (rows past the last tile: zero fills) must not outlive the CTA's shared memory
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    umma::mbar_wait(barR, phR);
  }

  mark(15);
  // ------------------------------------------------------------------ flush the gradient partial
  // Every epilogue is done and every GEMM waited for: the activation buffers are free.  The dW^T accumulators (TMEM lane =
  // output unit, columns = inputs) are transposed through shared memory so that the partial is written with coalesced
  // stores; Gv / Ga / Gm are turned into the gradients of the five layers behind the composites there.
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  float* part = a.partial + (size_t)blockIdx.x * a.partial_stride;
  const tc::LayerOffsets& lo = a.lo;
  const bool low = lane < 16;          // M = 64 accumulators: output / input unit mo at lanes 0-15 of the quadrant
  const int mo = q * 16 + (lane & 15);
  float* stg = reinterpret_cast<float*>(smem + OFF_XH1);  // XH1 | XH2 | XB1: 102 KB
  constexpr int S_ENC1 = 0, S_V1 = 64 * 65, S_GV = 2 * 64 * 65, S_GA = 3 * 64 * 65, S_ENC0 = 4 * 64 * 65, S_GM = S_ENC0 + 64 * 9;
  // fp32 weights of the layers behind the composites: W2^T [64][16] | b2 [16] | Wv0 [64][16] | Wa0 [64][16] | Wmap [8][16] |
  // Wa1 [16][65] (odd pitch) | ba1 [16]
  constexpr int S_W2T = S_GM + 16 * 65, S_B2 = S_W2T + 1024, S_WV0 = S_B2 + 16, S_WA0 = S_WV0 + 1024, S_WMAP = S_WA0 + 1024;
  constexpr int S_WA1 = S_WMAP + 128, S_BA1 = S_WA1 + 16 * 65, S_END = S_BA1 + 16;
  static_assert(S_END * 4 <= 51 * CG, "flush staging exceeds the XH1 | XH2 | XB1 region");
  // dW^T accumulators -> staging [out][pitch]; the bias gradient (column bias_col) goes straight out, or to staging column
  // `in_real` when part_bias < 0
  auto stage_T = [&](uint32_t col, int in_real, int in_pad, int bias_col, int part_bias, int soff, int pitch) {
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
    if (low) {
      if (part_bias >= 0) part[part_bias + mo] = v[0];
      else stg[soff + mo * pitch + in_real] = v[0];
    }
  };
  if (worker && g == 0) {
    stage_T(C_G0, a.obs_dim, 8, 8, lo.b[tc::L_ENC0], S_ENC0, 9);
    stage_T(C_GA, 64, 64, 64, -1, S_GA, 65);
    // Gm: TMEM lane = input unit, columns = logits -> staging [out][65]
    for (int c0 = 0; c0 < 16; c0 += 8) {
      float v[8];
      umma::tmem_ld8(tw + C_GM + c0, v);
      umma::tmem_ld_wait();
      if (low) {
#pragma unroll
        for (int j = 0; j < 8; ++j) stg[S_GM + (c0 + j) * 65 + mo] = v[j];
      }
    }
  }
  if (worker && g == 1) stage_T(C_G1, 64, 64, 64, lo.b[tc::L_ENC1], S_ENC1, 65);
  if (worker && g == 2) stage_T(C_GV1, 64, 64, 64, lo.b[tc::L_V1], S_V1, 65);
  if (worker && g == 3) stage_T(C_GV, 64, 64, 64, -1, S_GV, 65);
  for (int i = tid; i < 1024; i += THREADS) {
    const float w2 = __ldg(a.params + lo.w[tc::L_ENC2] + i);  // W2[k][i2]
    stg[S_W2T + (i & 63) * 16 + (i >> 6)] = w2;
    stg[S_WV0 + i] = __ldg(a.params + lo.w[tc::L_V0] + i);
    stg[S_WA0 + i] = __ldg(a.params + lo.w[tc::L_A0] + i);
    stg[S_WA1 + (i >> 6) * 65 + (i & 63)] = __ldg(a.params + lo.w[tc::L_A1] + i);
  }
  if (tid < 128) stg[S_WMAP + tid] = tid < n_out * 16 ? __ldg(a.params + lo.w[tc::L_MAP] + tid) : 0.f;
  if (tid >= 128 && tid < 144) stg[S_B2 + tid - 128] = __ldg(a.params + lo.b[tc::L_ENC2] + tid - 128);
  if (tid >= 160 && tid < 176) stg[S_BA1 + tid - 160] = __ldg(a.params + lo.b[tc::L_A1] + tid - 160);

  // ------------------------------------------------------------------ register-held gradients and loss statistics
  float* sh = reinterpret_cast<float*>(smem + OFF_DY64);
  float* s_wv2 = sh;                 // [4 quadrants][64 columns]
  float* s_bv2 = s_wv2 + 256;        // [4 quadrants]
  float* s_mb = s_bv2 + 4;           // [4 quadrants][LRX_MAX_NOUT]
  float* s_gmb = s_mb + 4 * LRX_MAX_NOUT;  // [LRX_MAX_NOUT]: column sums of d(logits) over the CTA's rows
  double* shd = reinterpret_cast<double*>(smem + OFF_DY64 + 4096);  // [6][16 warps]
  if (worker) {
    if ((lane & 1) == 0) s_wv2[q * 64 + cg + (lane >> 1)] = gw_v2;
    if (g0) {
      const float x = warp_sum(gb_v2);
      if (lane == 0) s_bv2[q] = x;
#pragma unroll
      for (int o2 = 0; o2 < LRX_MAX_NOUT; ++o2) {
        const float xb = o2 < NO ? warp_sum(gm_b[o2 < NO ? o2 : 0]) : 0.f;
        if (lane == 0) s_mb[q * LRX_MAX_NOUT + o2] = xb;
      }
    }
    double v6[6] = {(double)s_kl, (double)s_pol, (double)s_val, (double)s_ent, (double)s_dls, (double)s_bad};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      v6[k] = warp_sum(v6[k]);
      if (lane == 0) shd[k * 16 + warp] = v6[k];
    }
  }
  __syncthreads();
  auto sum4 = [](const float* p, int stride) { return ((p[0] + p[stride]) + p[2 * stride]) + p[3 * stride]; };
  if (tid < LRX_MAX_NOUT) s_gmb[tid] = sum4(s_mb + tid, LRX_MAX_NOUT);
  // ------------------------------------------------------------------ staging -> partial, coalesced; fixed-order sums
  {
    // 64 x 64 layers: shifts instead of a division by a run-time width; consecutive threads = consecutive addresses
    for (int e = tid; e < 2 * 4096; e += THREADS) {
      const int l2 = e >> 12, r = e & 4095, o = r >> 6, i = r & 63;
      part[lo.w[l2 ? tc::L_V1 : tc::L_ENC1] + r] = stg[(l2 ? S_V1 : S_ENC1) + o * 65 + i];
    }
    for (int e = tid; e < 64 * a.obs_dim; e += THREADS) {
      const int o = e / a.obs_dim, i = e - o * a.obs_dim;
      part[lo.w[tc::L_ENC0] + e] = stg[S_ENC0 + o * 9 + i];
    }
    if (tid < 64) part[lo.w[tc::L_V2] + tid] = sum4(s_wv2 + tid, 64);
    if (tid == 64) part[lo.b[tc::L_V2]] = sum4(s_bv2, 1);
    if (tid >= 320 && tid < 326) {
      const int k = tid - 320;
      double s2 = 0;
      for (int w = 0; w < 16; ++w) s2 += shd[k * 16 + w];
      a.loss_partial[(size_t)blockIdx.x * 8 + k] = s2;
    }
  }
  __syncthreads();  // s_gmb
  {
    // dWv0 / dWa0 [64][16] = G_h W2^T + g_1 b2^T, dbv0 / dba0 = g_1: warps 0-3 -> v0, 4-7 -> a0; thread = (output o, 8 of the 16 k)
    if (warp < 8) {
      const int head = warp >> 2, t = tid & 127, o = t & 63, kh = t >> 6;
      const float* G = stg + (head ? S_GA : S_GV) + o * 65;
      float acc[8];
      const float g1 = G[64];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll 8
      for (int i = 0; i < 64; ++i) {
        const float gi = G[i];
        const float4 w0 = *reinterpret_cast<const float4*>(stg + S_W2T + i * 16 + kh * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(stg + S_W2T + i * 16 + kh * 8 + 4);
        acc[0] = fmaf(gi, w0.x, acc[0]); acc[1] = fmaf(gi, w0.y, acc[1]); acc[2] = fmaf(gi, w0.z, acc[2]); acc[3] = fmaf(gi, w0.w, acc[3]);
        acc[4] = fmaf(gi, w1.x, acc[4]); acc[5] = fmaf(gi, w1.y, acc[5]); acc[6] = fmaf(gi, w1.z, acc[6]); acc[7] = fmaf(gi, w1.w, acc[7]);
      }
      float* dst = part + lo.w[head ? tc::L_A0 : tc::L_V0] + o * 16 + kh * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) dst[k] = fmaf(g1, stg[S_B2 + kh * 8 + k], acc[k]);
      if (kh == 0) part[lo.b[head ? tc::L_A0 : tc::L_V0] + o] = g1;
    } else if (warp < 16) {
      // dW2 [16][64] = Wv0^T Gv_h + Wa0^T Ga_h: thread = (input i, 4 of the 16 k)
      const int t = tid - 256, i = t & 63, kg = t >> 6;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
      for (int o = 0; o < 64; ++o) {
        const float gv = stg[S_GV + o * 65 + i], ga = stg[S_GA + o * 65 + i];
        const float4 wv = *reinterpret_cast<const float4*>(stg + S_WV0 + o * 16 + kg * 4);
        const float4 wa = *reinterpret_cast<const float4*>(stg + S_WA0 + o * 16 + kg * 4);
        acc[0] = fmaf(wa.x, ga, fmaf(wv.x, gv, acc[0])); acc[1] = fmaf(wa.y, ga, fmaf(wv.y, gv, acc[1]));
        acc[2] = fmaf(wa.z, ga, fmaf(wv.z, gv, acc[2])); acc[3] = fmaf(wa.w, ga, fmaf(wv.w, gv, acc[3]));
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) part[lo.w[tc::L_ENC2] + (kg * 4 + k) * 64 + i] = acc[k];
    } else {
      // the issuer warp: db2 [16] = Wv0^T gv_1 + Wa0^T ga_1 (lanes 0-15), dba1 [16] = Wmap^T gm_1 (lanes 16-31)
      const int k = lane & 15;
      float acc = 0.f;
      if (lane < 16) {
        for (int o = 0; o < 64; ++o)
          acc = fmaf(stg[S_WA0 + o * 16 + k], stg[S_GA + o * 65 + 64], fmaf(stg[S_WV0 + o * 16 + k], stg[S_GV + o * 65 + 64], acc));
        part[lo.b[tc::L_ENC2] + k] = acc;
      } else {
        for (int o = 0; o < n_out; ++o) acc = fmaf(stg[S_WMAP + o * 16 + k], s_gmb[o], acc);
        part[lo.b[tc::L_A1] + k] = acc;
      }
    }
    // dWa1 [16][64] = Wmap^T Gm^T: 1024 outputs, <= 8 terms each
    for (int e = tid; e < 1024; e += THREADS) {
      const int k = e >> 6, i = e & 63;
      float acc = 0.f;
      for (int o = 0; o < n_out; ++o) acc = fmaf(stg[S_WMAP + o * 16 + k], stg[S_GM + o * 65 + i], acc);
      part[lo.w[tc::L_A1] + e] = acc;
    }
    // dWmap [n_out][16] = Gm^T Wa1^T + gm_1 ba1^T, dbmap = gm_1
    if (tid < n_out * 16) {
      const int o = tid >> 4, k = tid & 15;
      float acc = 0.f;
      for (int i = 0; i < 64; ++i) acc = fmaf(stg[S_GM + o * 65 + i], stg[S_WA1 + k * 65 + i], acc);
      part[lo.w[tc::L_MAP] + tid] = fmaf(s_gmb[o], stg[S_BA1 + k], acc);
    }
    if (tid >= 256 && tid < 256 + n_out) part[lo.b[tc::L_MAP] + tid - 256] = s_gmb[tid - 256];
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
long long* lrx_tc128_prof_ptr();  // lrx_update_tc128.cu (test hook shared by both 128-row kernels)

// Same workspace layout as the other default-width kernels (lrx_fused_workspace_bytes covers all of them).
int lrx_ppo_grad_tc128c(const LrxPolicyDesc* pol, const float* params, const LrxBatchView* buf, const int32_t* idx,
                        int64_t B, int64_t B_global, const double* adv_sums, const LrxPpoHyper* h, float* gradbuf,
                        void* workspace, size_t workspace_bytes, cudaStream_t stream, bool* handled,
                        const LrxApplyArgs* apply) {
  *handled = false;
  if (!tc::is_default_policy(*pol) || lrx_fused_enabled() < 4) return LRX_OK;
  if (buf->packed != nullptr && (reinterpret_cast<uintptr_t>(buf->packed) & 15) != 0) return LRX_OK;  // 16-byte cp.async sources
  FusedWorkspace w = carve_fused(*pol, workspace);
  LRX_REQUIRE(workspace_bytes >= w.total, LRX_ERR_INVALID_ARG, "lrx_ppo_grad: workspace too small for the fused path");
  Tc128cArgs a{};
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
  a.prof = lrx_tc128_prof_ptr();
  const int sms = lrx_sm_count();
  const int grid = a.n_tiles < sms ? a.n_tiles : sms;
  static bool set0[64], set1[64], set2[64], set3[64];
  LRX_CUDA_CHECK(lrx_set_max_smem(ppo_grad_tc128c_kernel<0>, (int)SMEM_BYTES, set0));
  LRX_CUDA_CHECK(lrx_set_max_smem(ppo_grad_tc128c_kernel<1>, (int)SMEM_BYTES, set1));
  LRX_CUDA_CHECK(lrx_set_max_smem(ppo_grad_tc128c_kernel<2>, (int)SMEM_BYTES, set2));
  LRX_CUDA_CHECK(lrx_set_max_smem(ppo_grad_tc128c_kernel<3>, (int)SMEM_BYTES, set3));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = lrx_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  switch (pol->n_out) {
    case 1: LRX_CUDA_CHECK(cudaLaunchKernelEx(&cfg, ppo_grad_tc128c_kernel<1>, a)); break;
    case 2: LRX_CUDA_CHECK(cudaLaunchKernelEx(&cfg, ppo_grad_tc128c_kernel<2>, a)); break;
    case 3: LRX_CUDA_CHECK(cudaLaunchKernelEx(&cfg, ppo_grad_tc128c_kernel<3>, a)); break;
    default: LRX_CUDA_CHECK(cudaLaunchKernelEx(&cfg, ppo_grad_tc128c_kernel<0>, a)); break;
  }
  LRX_LAUNCH_CHECK();
  const int rc = lrx_launch_reduce_partials(w.partial, grid, (long long)w.P_pad, pol->n_params, w.loss_partial, grid,
                                            a.lo.log_std, h->entropy_loss_coefficient, (double)B_global, gradbuf, apply, stream);
  if (rc) return rc;
  LRX_LAUNCH_CHECK();
  *handled = true;
  return LRX_OK;
}
