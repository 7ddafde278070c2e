// Generic-width tcgen05 GEMM for the layer-by-layer PPO update (any MLPActorCriticPolicy widths; BASELINE.json
// configs[2]: Pendulum, 2-layer 256-wide MLPs).  Drop-in for sgemm_kernel of lrx_update.cu: same GemmArgs, same
// epilogue (bias, ReLU, ReLU mask, accumulate, ones row, redirected bias-gradient column, split-K).
//
//     C[M x N] (+)= op(A)[M x K] * op(B)[K x N]      fp32 in HBM, two-piece bf16 emulation on the tensor cores
//
// One CTA (512 threads, one per SM) = one 128 x (<= 288) accumulator in TMEM, walked over K in chunks of 32 through a
// two-deep asynchronous pipeline.  These GEMMs are HBM / latency bound (a 65 536-row minibatch of 256-wide activations
// is 67 MB per operand, nothing stays in L2), so the structure is built around keeping loads in flight:
//   1. cp.async (LDGSTS, zero-filling the ragged edges) copies the raw fp32 chunk k+1 / k+2 of both operands from
//      global memory straight into a shared-memory ring -- no registers, so 50 - 100 KB per SM are in flight while
//   2. all 16 warps convert chunk k from the ring to bf16 hi / lo (lrx_tc2.cuh) in the canonical no-swizzle layout,
//      which serves as a K-major operand when K is the contiguous axis of the source and as an MN-major operand when
//      the M/N index is -- so all four orientations of sgemm_kernel<AT, BT> feed tcgen05.mma without a transpose --
//   3. one elected thread issues the MMAs of chunk k (3 products hi*lo + lo*hi + hi*hi per 16-wide K step) and commits
//      them to the mbarrier that frees chunk k's operand buffer two chunks later.
// (The first version staged through registers, 4 x 32 bytes in flight per thread, and waited for every chunk's MMAs:
// ncu showed 40 - 55 % of the stall samples on the first use of a loaded value, 0.66 TB/s, tensor pipe 0.5 - 2 %; its
// dX epilogue read the ReLU mask one dependent load at a time.)
// SWAP decides which index of C rides on the 128 TMEM lanes:
//   SWAP = 1 (forward / dX: N = minibatch rows, M = layer width): lanes = j, columns = i -> C[i][j] is written with
//            consecutive lanes on consecutive addresses (coalesced);
//   SWAP = 0 (weight gradients: M = out, N = in + 1 <= 288 in ONE column tile, K = rows, split-K): lanes = i.
// The raw ring keeps 16-byte chunk p of row r at position p ^ (r & 7): both the LDGSTS writes (lanes along a row) and the
// converters' reads (a quarter-warp = 8 rows of one chunk column) are bank-conflict free.
#include "lrx_common.cuh"
#include "lrx_gemm.cuh"
#include "lrx_tc2.cuh"

namespace {

constexpr int KC = 32;          // K chunk
constexpr int UM = 128;         // UMMA M (TMEM lanes)
constexpr int UN_MAX = 288;     // accumulator columns per CTA: one MMA of N <= 256 plus one of N <= 32
constexpr int THREADS = 512;
constexpr uint32_t RAW_X = UM * KC * 4, RAW_Y = UN_MAX * KC * 4;              // fp32 ring, per stage
constexpr uint32_t CAN_X_PIECE = UM * KC * 2, CAN_Y_PIECE = UN_MAX * KC * 2;  // bf16 operand pieces, per stage
constexpr uint32_t STAGE_RAW = RAW_X + RAW_Y, STAGE_CAN = 2 * CAN_X_PIECE + 2 * CAN_Y_PIECE;
constexpr uint32_t OFF_RAW = 0, OFF_CAN = 2 * STAGE_RAW, OFF_MISC = OFF_CAN + 2 * STAGE_CAN;
constexpr uint32_t SMEM_BYTES = OFF_MISC + 64;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

// Source of one UMMA operand: element (mn, k) of op(.) lives at base[mn * s_mn + k * s_k]; exactly one stride is 1.
struct Src {
  const float* base;
  long long s_mn, s_k;
  int mn_total;
  int al16;  // every 4-element group along the contiguous axis starts on a 16-byte boundary
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 4 consecutive elements (`valid` of them inside the matrix, the rest zero filled) -> one 16-byte ring chunk
__device__ __forceinline__ void copy4(uint32_t dst, const float* src, int valid, int al16) {
  if (al16) {
    cp_async16(dst, src, valid * 4);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) cp_async4(dst + 4 * e, e < valid ? src + e : src, e < valid ? 4 : 0);
  }
}

__device__ __forceinline__ uint32_t ring_pitch(bool mn_contig, int mn_tile) {  // bytes per ring row, a multiple of 128
  return mn_contig ? (uint32_t)((mn_tile * 4 + 127) & ~127) : (uint32_t)(KC * 4);
}

// Everything about one thread's copies and conversions that does not depend on the K chunk is worked out ONCE per CTA (the
// chunk loop is instruction-issue bound: ~8 instructions per 16-byte copy and ~30 per converted unit is what is left).
constexpr int NXC = UM * (KC / 4) / THREADS;                         // 16-byte copies per thread and chunk, X operand
constexpr int NYC = (UN_MAX * (KC / 4) + THREADS - 1) / THREADS;     // ... Y operand (the last ones may be idle)
constexpr int NXU = UM * (KC / 8) / THREADS;                         // converted units per thread and chunk, X operand
constexpr int NYU = (UN_MAX * (KC / 8) + THREADS - 1) / THREADS;
static_assert(NXC * THREADS == UM * (KC / 4) && NXU * THREADS == UM * (KC / 8), "X work divides evenly");

struct CopyPlan {
  const float* src;  // first element of the 4-element group in chunk 0; advanced by one chunk per issue
  uint32_t dst;      // byte offset inside the operand's ring stage
  int vmn;           // valid elements along mn: -1 = no copy at all, 0 = zero fill (row outside the matrix)
  int kpos;          // k offset of the group inside a chunk (k contiguous: first of its 4 k; mn contiguous: its k row)
};

// 16-byte copy number `id` of rows [mn0, mn0 + mn_tile) x [kbeg + chunk * KC, ...) of one operand.
//   k contiguous : ring row = mn, KC floats     mn contiguous: ring row = k, mn_tile floats (pitch rounded up to 128 B)
template <bool MN_CONTIG>
__device__ __forceinline__ CopyPlan plan_copy(const Src& s, int mn0, int mn_tile, int kbeg, int id) {
  CopyPlan c;
  if (!MN_CONTIG) {
    const int r = id >> 3, p = id & 7, mn = mn0 + r;
    c.vmn = id < mn_tile * (KC / 4) ? (mn < s.mn_total ? 4 : 0) : -1;
    c.src = c.vmn > 0 ? s.base + (long long)mn * s.s_mn + kbeg + p * 4 : s.base;
    c.dst = (uint32_t)r * (KC * 4) + (uint32_t)((p ^ (r & 7)) << 4);
    c.kpos = p * 4;
  } else {
    const int cpr = mn_tile >> 2;  // 16-byte chunks per ring row
    const int r = id / cpr, p = id - r * cpr, mn = mn0 + p * 4;
    const int v = s.mn_total - mn;
    c.vmn = id < KC * cpr ? (v < 0 ? 0 : (v > 4 ? 4 : v)) : -1;
    c.src = c.vmn > 0 ? s.base + (long long)(kbeg + r) * s.s_k + mn : s.base;
    c.dst = (uint32_t)r * ring_pitch(true, mn_tile) + (uint32_t)((p ^ (r & 7)) << 4);
    c.kpos = r;
  }
  return c;
}

// Start copy `c` for the chunk whose first k is `krem` short of kend; `raw` = shared address of the operand's ring stage.
template <bool MN_CONTIG>
__device__ __forceinline__ void issue_copy(CopyPlan& c, const Src& s, uint32_t raw, int krem, long long kadv) {
  if (c.vmn < 0) return;
  int valid;
  if (!MN_CONTIG) {
    valid = krem - c.kpos;
    valid = c.vmn == 0 || valid < 0 ? 0 : (valid > 4 ? 4 : valid);
  } else {
    valid = c.kpos < krem ? c.vmn : 0;
  }
  const float* src = valid ? c.src : s.base;
  if (s.al16) {
    cp_async16(raw + c.dst, src, valid * 4);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) cp_async4(raw + c.dst + 4 * e, e < valid ? src + e : s.base, e < valid ? 4 : 0);
  }
  c.src += kadv;
}

struct UnitPlan {
  uint32_t src;  // ring offset of the first of the two 16-byte chunks (its partner is at src ^ 16)
  uint32_t dst;  // canonical offset of the unit in the hi piece; 0xFFFFFFFF = idle
};

// Unit = 8 elements along the contiguous axis.  A quarter-warp takes 8 ring rows of one unit column: conflict-free 16-byte
// reads through the XOR placement, and 8 consecutive 16-byte rows of one canonical chunk column on the way out.
//   k contiguous : matrix r = mn, c = k   (R = mn_tile rows)  -> K-major operand
//   mn contiguous: matrix r = k,  c = mn  (R = KC rows)       -> MN-major operand
template <bool MN_CONTIG>
__device__ __forceinline__ UnitPlan plan_unit(int mn_tile, int u) {
  UnitPlan p;
  const int chunks = MN_CONTIG ? mn_tile >> 3 : KC / 8;
  const int cb_shift = (chunks & 3) ? 1 : 2;
  const int r8 = u & 7, cl = (u >> 3) & ((1 << cb_shift) - 1), blk = u >> (3 + cb_shift);
  int r, ch;
  if (MN_CONTIG) {  // KC / 8 = 4 row blocks
    r = ((blk & 3) << 3) | r8;
    ch = ((blk >> 2) << cb_shift) | cl;
  } else {          // 4 unit columns per row: blk = row block
    r = (blk << 3) | r8;
    ch = cl;
  }
  const uint32_t R = MN_CONTIG ? KC : mn_tile;
  p.src = (uint32_t)r * ring_pitch(MN_CONTIG, mn_tile) + (uint32_t)(((2 * ch) ^ r8) << 4);
  p.dst = u < mn_tile * (KC / 8) ? (uint32_t)ch * (R * 16u) + (uint32_t)r * 16u : 0xFFFFFFFFu;
  return p;
}

__device__ __forceinline__ void convert_unit(const UnitPlan& p, const uint8_t* raw, uint8_t* can, uint32_t lo_off) {
  if (p.dst == 0xFFFFFFFFu) return;
  const float4 a = *reinterpret_cast<const float4*>(raw + p.src);
  const float4 b = *reinterpret_cast<const float4*>(raw + (p.src ^ 16u));
  uint32_t h[4], l[4];
  tc2::split2(a.x, a.y, h[0], l[0]);
  tc2::split2(a.z, a.w, h[1], l[1]);
  tc2::split2(b.x, b.y, h[2], l[2]);
  tc2::split2(b.z, b.w, h[3], l[3]);
  *reinterpret_cast<uint4*>(can + p.dst) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(can + p.dst + lo_off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// X_MN / Y_MN: the UMMA A / B operand's source has the MN index contiguous (-> MN-major descriptor).
// persistent != 0 (one column tile, <= 256 columns): gridDim.x CTAs walk the lane tiles blockIdx.x, + gridDim.x, ... as
// ONE continuous stream of K chunks -- the copies of the next tile's first chunks are in flight and its first MMAs
// issued (into the other half of TMEM) before the previous tile's epilogue runs, so neither the DRAM latency at a
// tile's start nor the commit latency at its end is exposed.  Otherwise one tile per CTA (grid = tiles).
template <bool SWAP, bool X_MN, bool Y_MN>
__global__ void __launch_bounds__(THREADS, 1)
tc_gemm_kernel(const GemmArgs g, const int x_al16, const int y_al16, const int persistent) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // programmatic dependent launch: a layer-by-layer update is 30 dependent launches per minibatch; this grid may be set up
  // while the previous GEMM drains (nothing of its output is read before the wait)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int tid = threadIdx.x, warp = umma::warp_idx_sync(), lane = tid & 31;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_MISC);  // [2]: the MMAs that read operand stage s are done
  uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + OFF_MISC + 16);

  // tile of C: x -> the index on the TMEM lanes (128 per tile), y -> the index on the TMEM columns (<= 288 per tile)
  const int x_total = SWAP ? g.N : g.M, y_total = SWAP ? g.M : g.N;
  const int x_tiles = (x_total + UM - 1) / UM;
  const int G = persistent ? (int)gridDim.x : x_tiles;
  const int my_tiles = persistent ? (x_tiles - (int)blockIdx.x + G - 1) / G : 1;
  auto tile_x0 = [&](int t) { return ((int)blockIdx.x + t * G) * UM; };
  const int y0 = blockIdx.y * UN_MAX;
  const int y_cnt = min(UN_MAX, y_total - y0);
  const int un = (y_cnt + 15) & ~15;  // UMMA N: multiple of 16 for M = 128
  const int n1 = min(un, 256), n2 = un - n1;
  const int kbeg = blockIdx.z * g.k_chunk;
  const int kend = min(g.K, kbeg + g.k_chunk);
  const int nk = kend > kbeg ? (kend - kbeg + KC - 1) / KC : 0;
  const int total = my_tiles * nk;  // chunks this CTA streams

  // op(A)[i][k]: AT ? A[k * lda + i] : A[i * lda + k];   op(B)[k][j]: BT ? B[j * ldb + k] : B[k * ldb + j]
  // X carries the lane index (SWAP: j from B, else i from A), Y the column index.
  Src sx, sy;
  if (SWAP) {
    sx = Src{g.B, X_MN ? 1 : g.ldb, X_MN ? g.ldb : 1, g.N, x_al16};  // mn = j
    sy = Src{g.A, Y_MN ? 1 : g.lda, Y_MN ? g.lda : 1, g.M, y_al16};  // mn = i
  } else {
    sx = Src{g.A, X_MN ? 1 : g.lda, X_MN ? g.lda : 1, g.M, x_al16};
    sy = Src{g.B, Y_MN ? 1 : g.ldb, Y_MN ? g.ldb : 1, g.N, y_al16};
  }
  CopyPlan cx[NXC], cy[NYC];
  UnitPlan ux[NXU], uy[NYU];
#pragma unroll
  for (int n = 0; n < NXU; ++n) ux[n] = plan_unit<X_MN>(UM, tid + n * THREADS);
#pragma unroll
  for (int n = 0; n < NYU; ++n) uy[n] = plan_unit<Y_MN>(un, tid + n * THREADS);
  const long long kadv_x = X_MN ? (long long)KC * sx.s_k : KC, kadv_y = Y_MN ? (long long)KC * sy.s_k : KC;
  const uint32_t smem_base = umma::smem_u32(smem);
  const bool al_both = sx.al16 && sy.al16;
  int i_tile = 0, i_kc = 0, i_gc = 0;  // the next chunk to issue: tile, chunk inside the tile, position in the stream
  auto issue_next = [&]() {            // -> ring stage i_gc & 1
    if (i_kc == 0) {                   // a new tile: its copies start over at kbeg
#pragma unroll
      for (int n = 0; n < NXC; ++n) cx[n] = plan_copy<X_MN>(sx, tile_x0(i_tile), UM, kbeg, tid + n * THREADS);
#pragma unroll
      for (int n = 0; n < NYC; ++n) cy[n] = plan_copy<Y_MN>(sy, y0, un, kbeg, tid + n * THREADS);
    }
    const uint32_t raw = smem_base + OFF_RAW + (uint32_t)(i_gc & 1) * STAGE_RAW;
    const int krem = kend - (kbeg + i_kc * KC);
    if (al_both && krem >= KC) {  // the common case: one LDGSTS.128 and a pointer bump per copy
#pragma unroll
      for (int n = 0; n < NXC; ++n) {
        if (cx[n].vmn >= 0) cp_async16(raw + cx[n].dst, cx[n].vmn ? cx[n].src : sx.base, cx[n].vmn * 4);
        cx[n].src += kadv_x;
      }
#pragma unroll
      for (int n = 0; n < NYC; ++n) {
        if (cy[n].vmn >= 0) cp_async16(raw + RAW_X + cy[n].dst, cy[n].vmn ? cy[n].src : sy.base, cy[n].vmn * 4);
        cy[n].src += kadv_y;
      }
    } else {
#pragma unroll
      for (int n = 0; n < NXC; ++n) issue_copy<X_MN>(cx[n], sx, raw, krem, kadv_x);
#pragma unroll
      for (int n = 0; n < NYC; ++n) issue_copy<Y_MN>(cy[n], sy, raw + RAW_X, krem, kadv_y);
    }
    ++i_gc;
    if (++i_kc == nk) {
      i_kc = 0;
      ++i_tile;
    }
  };

  if (tid == 0) {
    umma::mbar_init(bar, 1);
    umma::mbar_init(bar + 1, 1);
    umma::fence_mbar_init();
  }
  if (warp == 0) umma::tmem_alloc(tslot, 512);
  if (total > 0) issue_next();
  cp_async_commit();
  if (total > 1) issue_next();
  cp_async_commit();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tb = *tslot;

  // ---- epilogue of tile e: warp w reads TMEM quadrant w & 3 (lanes = x), 16 of every 64 columns.  Element t of a group
  // is C[i][j] with (i, j) = SWAP ? (yb + t, x) : (x, yb + t): base pointer + t * stride.  Warp-level only (no CTA barrier).
  const int q = warp & 3, cgp = warp >> 2;
  float* C = g.C + (size_t)blockIdx.z * g.c_split_stride;
  float* C2 = g.C2 ? g.C2 + (size_t)blockIdx.z * g.c_split_stride : nullptr;
  const int cstride = SWAP ? g.ldc : 1, mstride = SWAP ? g.ldm : 1;
  const bool relu = g.relu != 0;
  auto epilogue = [&](int e) {
    const int x0 = tile_x0(e);
    const int x = x0 + q * 32 + lane;
    const bool x_ok = x < x_total;
    const uint32_t tw = umma::tmem_addr(tb, q * 32, persistent ? (uint32_t)(e & 1) * 256u : 0u);
    const float bias_x = (!SWAP && g.bias && x_ok) ? __ldg(g.bias + x) : 0.f;
    for (int c0 = cgp * 16; c0 < un; c0 += 64) {
      float v[16];
      if (nk > 0) umma::tmem_ld16(tw + c0, v);
      const int yb = y0 + c0;
      const int ny = min(16, y_total - yb);  // valid elements of the group (<= 0: padding columns only)
      float* cp = C + (SWAP ? (size_t)yb * g.ldc + x : (size_t)x * g.ldc + yb);
      const float* mp = g.mask ? g.mask + (SWAP ? (size_t)yb * g.ldm + x : (size_t)x * g.ldm + yb) : nullptr;
      // warp-uniform choice: tcgen05.wait::ld below is .sync.aligned, the whole warp must reach the same one
      if (__all_sync(0xffffffffu, x_ok) && ny == 16 && !C2) {  // interior group, no redirected column: straight-line
        // every global read of the group is issued before the first use and before the TMEM wait (the mask and the
        // accumulate target may alias C for all the compiler knows: interleaved with the stores they would be 16
        // dependent round trips to HBM)
        float bv[16], mk[16], old[16];
        const float* bp = (SWAP && g.bias) ? g.bias + yb : nullptr;
#pragma unroll
        for (int t = 0; t < 16; ++t) bv[t] = bp ? __ldg(bp + t) : bias_x;
        if (mp) {
#pragma unroll
          for (int t = 0; t < 16; ++t) mk[t] = __ldg(mp + t * mstride);
        }
        if (g.accumulate) {
#pragma unroll
          for (int t = 0; t < 16; ++t) old[t] = cp[t * cstride];
        }
        if (nk > 0) {
          umma::tmem_ld_wait();
        } else {
#pragma unroll
          for (int t = 0; t < 16; ++t) v[t] = 0.f;
        }
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const float r = v[t] + bv[t];
          v[t] = relu ? fmaxf(r, 0.f) : r;
        }
        if (mp) {
#pragma unroll
          for (int t = 0; t < 16; ++t) v[t] = mk[t] > 0.f ? v[t] : 0.f;
        }
        if (g.accumulate) {
#pragma unroll
          for (int t = 0; t < 16; ++t) v[t] += old[t];
        }
#pragma unroll
        for (int t = 0; t < 16; ++t) cp[t * cstride] = v[t];
      } else {
        if (nk > 0) {
          umma::tmem_ld_wait();
        } else {
#pragma unroll
          for (int t = 0; t < 16; ++t) v[t] = 0.f;
        }
        if (x_ok) {
#pragma unroll 1
          for (int t = 0; t < ny; ++t) {
            const int i = SWAP ? yb + t : x, j = SWAP ? x : yb + t;
            float* dst = (C2 && j == g.n_main) ? (C2 + i) : (cp + (size_t)t * cstride);
            float r = v[0] + (g.bias ? __ldg(g.bias + i) : 0.f);
            if (relu) r = fmaxf(r, 0.f);
            if (mp) r = __ldg(mp + (size_t)t * mstride) > 0.f ? r : 0.f;
            if (g.accumulate) r += *dst;
            *dst = r;
#pragma unroll
            for (int e2 = 0; e2 < 15; ++e2) v[e2] = v[e2 + 1];  // rotate: keeps v[] in registers under the rolled loop
          }
        }
      }
    }
    if (g.ones_row && blockIdx.z == 0 && (SWAP ? blockIdx.y == 0 : x0 == 0)) {
      // C[M][j] = 1 for the j of this tile
      const int jn = SWAP ? UM : y_cnt, jb = SWAP ? x0 : y0;
      for (int t = tid; t < jn; t += THREADS)
        if (jb + t < g.N) g.C[(size_t)g.M * g.ldc + jb + t] = 1.0f;
    }
  };

  uint32_t phase[2] = {0u, 0u};
  int t_cur = 0, kc = 0, epi = 0;  // tile / chunk of stream position gc; the next tile whose epilogue is due
  for (int gc = 0; gc < total; ++gc) {
    const int st = gc & 1;
    uint8_t* can = smem + OFF_CAN + (uint32_t)st * STAGE_CAN;
    const tc2::Mat X{can, (uint32_t)(X_MN ? KC : UM), CAN_X_PIECE};
    const tc2::Mat Y{can + 2 * CAN_X_PIECE, (uint32_t)(Y_MN ? KC : un), CAN_Y_PIECE};
    cp_async_wait<1>();  // this thread's copies of chunk gc have landed (chunk gc + 1 may still be in flight)
    if (gc >= 2) {       // the MMAs of chunk gc - 2 have read operand stage st
      umma::mbar_wait(bar + st, phase[st]);
      phase[st] ^= 1u;
      umma::tc_fence_after();
    }
    __syncthreads();     // every thread's copies of chunk gc are visible
    // tile `epi` finished with chunk (epi + 1) * nk - 1, which the wait above has covered two chunks later: its epilogue
    // runs here, under the copies of chunks gc, gc + 1 and with the next tile's first MMAs already queued or done
    if (epi < t_cur && gc >= (epi + 1) * nk + 1) {
      epilogue(epi);
      ++epi;
    }
    const uint8_t* raw = smem + OFF_RAW + (uint32_t)st * STAGE_RAW;
#pragma unroll
    for (int n = 0; n < NXU; ++n) convert_unit(ux[n], raw, X.ptr, CAN_X_PIECE);
#pragma unroll
    for (int n = 0; n < NYU; ++n) convert_unit(uy[n], raw + RAW_X, Y.ptr, CAN_Y_PIECE);
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();     // operand stage complete, ring stage st free, this iteration's epilogue reads of TMEM done
    if (gc + 2 < total) issue_next();
    cp_async_commit();
    if (warp == 0) {
      umma::tc_fence_after();
      const bool leader = umma::elect_one();
      const int ksteps = (min(KC, kend - (kbeg + kc * KC)) + 15) >> 4;  // zero-filled beyond kend
      const uint32_t d = tb + (persistent ? (uint32_t)(t_cur & 1) * 256u : 0u);
      tc2::gemm<X_MN, Y_MN>(leader, d, X, Y, UM, n1, n1, ksteps, kc > 0 ? 1u : 0u);
      if (n2) {
        const tc2::Mat Y2{Y.ptr + (Y_MN ? 32u * (KC * 16u) : 256u * 16u), Y.R, CAN_Y_PIECE};  // columns 256 ..
        tc2::gemm<X_MN, Y_MN>(leader, d + 256, X, Y2, UM, n2, n2, ksteps, kc > 0 ? 1u : 0u);
      }
      if (leader) umma::mma_commit(bar + st);
      __syncwarp();
    }
    if (++kc == nk) {
      kc = 0;
      ++t_cur;
    }
  }
  cp_async_wait<0>();
  if (total > 0) {  // the last commit covers every earlier MMA
    umma::mbar_wait(bar + ((total - 1) & 1), phase[(total - 1) & 1]);
    umma::tc_fence_after();
  }
  while (epi < my_tiles) epilogue(epi++);
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tb, 512);
}

static int aligned16(const float* base, long long stride_of_rows) {
  return ((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (stride_of_rows & 3) == 0) ? 1 : 0;
}

template <bool SWAP, bool X_MN, bool Y_MN>
int launch(const GemmArgs& g, int splits, cudaStream_t s) {
  static bool attr_set[64];
  LRX_CUDA_CHECK(lrx_set_max_smem(tc_gemm_kernel<SWAP, X_MN, Y_MN>, (int)SMEM_BYTES, attr_set));
  const int x_total = SWAP ? g.N : g.M, y_total = SWAP ? g.M : g.N;
  // the non-unit stride of each operand's source (rows of the ring are that far apart in global memory)
  const float* xb = SWAP ? g.B : g.A;
  const float* yb = SWAP ? g.A : g.B;
  const int xl = SWAP ? g.ldb : g.lda, yl = SWAP ? g.lda : g.ldb;
  const int x_tiles = (x_total + UM - 1) / UM, y_tiles = (y_total + UN_MAX - 1) / UN_MAX;
  // one column tile of <= 256 columns: both halves of TMEM hold an accumulator and the CTAs stream over the lane tiles
  const int persistent = (y_tiles == 1 && y_total <= 256 && x_tiles > 1) ? 1 : 0;
  int gx = x_tiles;
  if (persistent) {
    const int slots = lrx_sm_count() / (splits > 0 ? splits : 1);
    gx = x_tiles < slots ? x_tiles : (slots > 0 ? slots : 1);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(gx, y_tiles, splits);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = lrx_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LRX_CUDA_CHECK(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<SWAP, X_MN, Y_MN>, g, aligned16(xb, xl), aligned16(yb, yl), persistent));
  return LRX_OK;
}

}  // namespace

// Returns LRX_OK and sets *handled when the shape is worth the tensor cores (both the contraction and the layer width at
// least 16; the d <= 8 first layer and the scalar heads stay on the FMA kernel).
int lrx_launch_gemm_tc(const GemmArgs& g, bool at, bool bt, int splits, cudaStream_t s, bool* handled) {
  *handled = false;
  if (lrx_fused_enabled() < 1) return LRX_OK;
  // narrow layers (the d <= 8 first layer, scalar heads) ride along zero padded to the 16-wide MMA minimum: their cost is
  // the HBM pass over the wide operand, which the FMA kernel made at a sixth of the bandwidth.  Tiny problems stay on it.
  const long long big = (long long)(g.M > g.N ? g.M : g.N) * (g.K > 16 ? g.K : 16);
  if (g.M < 1 || g.N < 1 || g.K < 1 || big < 128 * 64) return LRX_OK;
  // forward / dX: the minibatch rows are N and C's contiguous axis, so they ride on the lanes (coalesced stores).  Weight
  // gradients (the rows are K, C is small): whichever orientation stages fewer operand rows -- out x (in + 1) with
  // in + 1 <= 288 is one column tile per 128 outputs, not a third lane tile for the single bias column.
  bool rows_on_n = g.N >= g.M;
  if (g.K > g.M && g.K > g.N) {
    auto cost = [](int x, int y) {
      const long long xt = (x + UM - 1) / UM, yt = (y + UN_MAX - 1) / UN_MAX;
      return xt * yt * (UM + (y < UN_MAX ? y : UN_MAX));
    };
    rows_on_n = cost(g.N, g.M) < cost(g.M, g.N) || (cost(g.N, g.M) == cost(g.M, g.N) && g.N >= g.M);
  }
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
