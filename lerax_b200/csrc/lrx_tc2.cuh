// Two-piece ("bf16x2") fp32 emulation on tcgen05 for the PPO update kernel (lrx_update_tc128.cu).
//
// Every fp32 value x is stored as hi + lo, both bf16 by ROUND-TO-NEAREST: hi = bf16(x), lo = bf16(x - hi), so
// |x - hi - lo| <= 2^-18 |x|.  A product keeps a_hi*b_hi + a_hi*b_lo + a_lo*b_hi (the dropped a_lo*b_lo is
// <= 2^-18 relative), each an exact bf16 x bf16 product accumulated in fp32 in TMEM: ~1e-5 worst case, ~3e-6 rms of
// sum|a||b| per dot product.  The update's contract is 1e-4 on gradients and parameters (BASELINE.json north_star);
// the rollout, whose contract is 1e-5 on values and log-probs, keeps the exact three-piece form of lrx_tc.cuh.
// Half the tensor-core work and two thirds of the shared memory of the three-piece form: that is what lets a tile be
// 128 rows (UMMA M = 128) instead of 64.
//
// Operand layout = lrx_tc.cuh's canonical no-swizzle layout (2-byte elements):
//     element (r, c) of an R-row matrix  ->  (c / 8) * (R * 16) + r * 16 + (c % 8) * 2
// readable K-major (MN = r, K = c: LBO = R*16, SBO = 128) and MN-major (MN = c, K = r: LBO = 128, SBO = R*16).
// The lo piece starts `lo_off` bytes after the hi piece.  A matrix may carry ONE extra c-group in its hi piece
// holding the constant column [1, 0, ..., 0]: as the B operand of a weight-gradient GEMM it makes the bias gradient
// (the column sums of dY) one more accumulator column instead of a separate reduction.
#pragma once

#include <cuda_bf16.h>

#include "lrx_tc.cuh"

namespace tc2 {

struct Mat {
  uint8_t* ptr;     // hi piece
  uint32_t R;       // rows
  uint32_t lo_off;  // byte distance hi -> lo piece
  uint32_t cgs = 0; // byte distance between c-groups when it is not R * 16 (0 = contiguous): a two-c-group matrix may keep
                    // its second c-group anywhere, e.g. share another matrix's constant ones column
};
__device__ __forceinline__ uint32_t cg_stride(const Mat& m) { return m.cgs ? m.cgs : m.R * 16u; }

// two adjacent columns (c even at the lower address) -> hi word, lo word.
// The lo piece is stored NEGATED (lo' = bf16(hi - x)): the mixed-precision subtract of sm_100 (sub.f32.bf16, SASS FHADD.BF16
// with a half selector) takes the packed hi word as it is, so the split is four instructions per pair instead of six, and
// the two cross products of gemm() put the sign back with the instruction descriptor's negate bits.
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  uint32_t h, l;
  float da, db;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));
  asm("{\n\t.reg .b16 l16, u16;\n\tmov.b32 {l16, u16}, %2;\n\tsub.rn.f32.bf16 %0, l16, %3;\n\tsub.rn.f32.bf16 %1, u16, %4;\n\t}"
      : "=f"(da), "=f"(db)
      : "r"(h), "f"(a), "f"(b));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(db), "f"(da));
  hi = h;
  lo = l;
}
constexpr uint32_t NEG_A = 1u << 13, NEG_B = 1u << 14;  // instruction descriptor: negate the A / B operand

// 8 consecutive columns (one 16-byte chunk) of row r
__device__ __forceinline__ void store8(const Mat& m, int chunk, int r, const float* v) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split2(v[2 * i], v[2 * i + 1], h[i], l[i]);
  uint8_t* d = m.ptr + (uint32_t)chunk * (m.R * 16u) + (uint32_t)r * 16u;
  *reinterpret_cast<uint4*>(d) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(d + m.lo_off) = make_uint4(l[0], l[1], l[2], l[3]);
}
// 4 consecutive columns c0 .. c0+3 (c0 % 4 == 0) of row r
__device__ __forceinline__ void store4(const Mat& m, int c0, int r, const float* v) {
  uint32_t h[2], l[2];
  split2(v[0], v[1], h[0], l[0]);
  split2(v[2], v[3], h[1], l[1]);
  uint8_t* d = m.ptr + (uint32_t)(c0 >> 3) * (m.R * 16u) + (uint32_t)r * 16u + (uint32_t)(c0 & 7) * 2u;
  *reinterpret_cast<uint2*>(d) = make_uint2(h[0], h[1]);
  *reinterpret_cast<uint2*>(d + m.lo_off) = make_uint2(l[0], l[1]);
}

template <bool MN>
__device__ __forceinline__ uint32_t desc_hi(const Mat& m) {
  return ((MN ? cg_stride(m) : 128u) >> 4) | (1u << 14);  // SBO | version
}
template <bool MN>
__device__ __forceinline__ uint32_t desc_lo(const Mat& m) {
  return ((umma::smem_u32(m.ptr) >> 4) & 0x3FFFu) | (((MN ? 128u : cg_stride(m)) >> 4) << 16);  // addr | LBO
}
template <bool MN>
__device__ __forceinline__ uint32_t desc_kstep(const Mat& m) {
  return (MN ? 256u : 2u * cg_stride(m)) >> 4;
}

// D[tmem] (+)= A * B^T over `ksteps` slices of K = 16, three products per slice, smallest first:
//   A_hi B_lo (N_lo columns), A_lo B_hi (N columns), A_hi B_hi (N columns).
// N_lo < N only for a B operand with the extra ones column (whose lo piece does not exist); such a GEMM must
// accumulate (the first product would not initialise columns N_lo .. N).
// Called by a whole CONVERGED warp; `leader` (umma::elect_one()) predicates the instructions.
template <bool A_MN, bool B_MN>
__device__ __forceinline__ void gemm(bool leader, uint32_t d_tmem, const Mat& A, const Mat& B, int M, int N, int N_lo,
                                     int ksteps, uint32_t accumulate) {
  const uint32_t idesc = umma::idesc_bf16(M, N, A_MN ? 1 : 0, B_MN ? 1 : 0);
  const uint32_t idesc_lo = umma::idesc_bf16(M, N_lo, A_MN ? 1 : 0, B_MN ? 1 : 0);
  uint64_t ah = tc::make_desc(desc_lo<A_MN>(A), desc_hi<A_MN>(A));
  uint64_t al = tc::make_desc(desc_lo<A_MN>(A) + (A.lo_off >> 4), desc_hi<A_MN>(A));
  uint64_t bh = tc::make_desc(desc_lo<B_MN>(B), desc_hi<B_MN>(B));
  uint64_t bl = tc::make_desc(desc_lo<B_MN>(B) + (B.lo_off >> 4), desc_hi<B_MN>(B));
  const uint64_t aks = desc_kstep<A_MN>(A), bks = desc_kstep<B_MN>(B);
  uint32_t acc = accumulate;
#pragma unroll 1
  for (int ks = 0; ks < ksteps; ++ks) {
    if (leader) {
      umma::mma_f16(d_tmem, ah, bl, idesc_lo | NEG_B, acc);
      umma::mma_f16(d_tmem, al, bh, idesc | NEG_A, N_lo == N ? 1u : acc);
      umma::mma_f16(d_tmem, ah, bh, idesc, 1u);
    }
    acc = 1u;
    ah += aks; al += aks; bh += bks; bl += bks;  // start-address field only; never carries out of it
  }
}

// Stage a Linear's weight W[out][in] (fp32, row-major, global) as the canonical matrix r = out, c = in (zero-padded
// to in_pad, a multiple of 8).
__device__ __forceinline__ void stage_weight(const float* __restrict__ W, int out, int in, int in_pad, const Mat& dst,
                                             int tid, int nthreads) {
  const int chunks = in_pad >> 3;
  for (int e = tid; e < out * chunks; e += nthreads) {
    const int o = e % out, ch = e / out;
    float v[8];
    const float* src = W + (size_t)o * in + ch * 8;
    if (ch * 8 + 8 <= in && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = ch * 8 + j;
        v[j] = i < in ? __ldg(src + j) : 0.f;
      }
    }
    store8(dst, ch, o, v);
  }
}

}  // namespace tc2
