// Shared pieces of the fused tensor-core kernels (rollout and PPO update) for the default
// MLPActorCriticPolicy widths: fp32 emulation on tcgen05 with bf16 pieces ("bf16x3").
//
// Every fp32 value x is stored as three bf16 pieces x = p0 + p1 + p2 (8 mantissa bits each,
// obtained by truncation, so the decomposition is EXACT for any finite fp32).  A product of
// two such values keeps the six piece products with i + j <= 2 (p_i * q_j), each an exact
// bf16 x bf16 -> fp32 tensor-core product accumulated in fp32 in TMEM; the dropped terms are
// <= 2^-24 relative.  Measured against float64 on B200 (tests/test_gpu_umma.py): ~1.5e-7 of
// sum|a||b| at K = 16, ~7e-7 at K = 64 -- the same as an fp32 FMA chain of that length.
//
// Operands sit in shared memory in one canonical no-swizzle layout (2-byte elements):
//     element (r, c) of an R-row matrix  ->  (c / 8) * (R * 16) + r * 16 + (c % 8) * 2
// i.e. 16-byte chunks of 8 consecutive c, the R chunks of one c-group contiguous.  The same
// bytes are a K-major operand (MN = r, K = c: LBO = R*16, SBO = 128) and an MN-major operand
// (MN = c, K = r: LBO = 128, SBO = R*16), so activations written once by the forward pass
// feed the forward GEMM, the dX GEMM and the weight-gradient GEMM (contraction over rows)
// without a transpose.  (For 4-byte tf32 operands the MN-major form only exists in the
// 128B-swizzle/32B-atom layout, which no K-major form matches -- measured with
// tools/umma_probe*.py -- hence bf16 pieces rather than 3xTF32.)
#pragma once

#include "lrx_umma.cuh"

namespace tc {

// (bf16(lo) | bf16(hi) << 16), both by truncation: the upper halves of the two floats.
__device__ __forceinline__ uint32_t pack_hi16(float lo, float hi) {
  return __byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x7632);
}
__device__ __forceinline__ float trunc_bf16(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFF0000u); }

// 8 consecutive values -> one 16-byte chunk per piece.
__device__ __forceinline__ void split8(const float* v, uint4& p0, uint4& p1, uint4& p2) {
  float r1[8], r2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    r1[i] = v[i] - trunc_bf16(v[i]);
    r2[i] = r1[i] - trunc_bf16(r1[i]);
  }
  p0 = make_uint4(pack_hi16(v[0], v[1]), pack_hi16(v[2], v[3]), pack_hi16(v[4], v[5]), pack_hi16(v[6], v[7]));
  p1 = make_uint4(pack_hi16(r1[0], r1[1]), pack_hi16(r1[2], r1[3]), pack_hi16(r1[4], r1[5]), pack_hi16(r1[6], r1[7]));
  p2 = make_uint4(pack_hi16(r2[0], r2[1]), pack_hi16(r2[2], r2[3]), pack_hi16(r2[4], r2[5]), pack_hi16(r2[6], r2[7]));
}

// A matrix in the canonical layout: generic pointer / shared address of piece 0, rows R,
// byte distance between pieces.
struct Mat {
  uint8_t* ptr;
  uint32_t R;
  uint32_t piece_bytes;
  __device__ __forceinline__ uint32_t saddr(int piece) const { return umma::smem_u32(ptr) + piece * piece_bytes; }
  // Operand starting at MN index mn0 (multiple of 8): K-major view -> rows r >= mn0;
  // MN-major view -> columns c >= mn0.
  __device__ __forceinline__ Mat rows_from(uint32_t r0) const { return Mat{ptr + r0 * 16u, R, piece_bytes}; }
  __device__ __forceinline__ Mat cols_from(uint32_t c0) const { return Mat{ptr + (c0 >> 3) * (R * 16u), R, piece_bytes}; }
};

// Store 8 consecutive values (columns 8*chunk .. +7 of row r) as the three pieces.
__device__ __forceinline__ void store8(const Mat& m, int chunk, int r, const float* v) {
  uint4 p0, p1, p2;
  split8(v, p0, p1, p2);
  uint8_t* d = m.ptr + (uint32_t)chunk * (m.R * 16u) + (uint32_t)r * 16u;
  *reinterpret_cast<uint4*>(d) = p0;
  *reinterpret_cast<uint4*>(d + m.piece_bytes) = p1;
  *reinterpret_cast<uint4*>(d + 2 * m.piece_bytes) = p2;
}

__device__ __forceinline__ void store_pieces(const Mat& m, int chunk, int r, const uint4& p0, const uint4& p1,
                                             const uint4& p2) {
  uint8_t* d = m.ptr + (uint32_t)chunk * (m.R * 16u) + (uint32_t)r * 16u;
  *reinterpret_cast<uint4*>(d) = p0;
  *reinterpret_cast<uint4*>(d + m.piece_bytes) = p1;
  *reinterpret_cast<uint4*>(d + 2 * m.piece_bytes) = p2;
}

// Two adjacent columns (c, c+1; c even) of one row as the three piece words.
__device__ __forceinline__ void split2(float a, float b, uint32_t& w0, uint32_t& w1, uint32_t& w2) {
  const float ra = a - trunc_bf16(a), rb = b - trunc_bf16(b);
  const float sa = ra - trunc_bf16(ra), sb = rb - trunc_bf16(rb);
  w0 = pack_hi16(a, b);
  w1 = pack_hi16(ra, rb);
  w2 = pack_hi16(sa, sb);
}
__device__ __forceinline__ void store2(const Mat& m, int r, int c, uint32_t w0, uint32_t w1, uint32_t w2) {
  uint8_t* d = m.ptr + (uint32_t)(c >> 3) * (m.R * 16u) + (uint32_t)r * 16u + (uint32_t)(c & 7) * 2u;
  *reinterpret_cast<uint32_t*>(d) = w0;
  *reinterpret_cast<uint32_t*>(d + m.piece_bytes) = w1;
  *reinterpret_cast<uint32_t*>(d + 2 * m.piece_bytes) = w2;
}

// Descriptor halves of an operand view.  K-major view: K runs along c (2 chunks per K = 16 MMA);
// MN-major view: K runs along r (16 rows per MMA).  Only the start-address field (low word,
// 16-byte units) changes between pieces and K slices, so the issue loop is integer adds.
template <bool MN>
__device__ __forceinline__ uint32_t desc_hi(const Mat& m) {
  return ((MN ? m.R * 16u : 128u) >> 4) | (1u << 14);  // SBO | version
}
template <bool MN>
__device__ __forceinline__ uint32_t desc_lo(const Mat& m) {
  return ((umma::smem_u32(m.ptr) >> 4) & 0x3FFFu) | (((MN ? 128u : m.R * 16u) >> 4) << 16);  // addr | LBO
}
template <bool MN>
__device__ __forceinline__ uint32_t desc_kstep(const Mat& m) {
  return (MN ? 256u : 2u * m.R * 16u) >> 4;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// D[tmem] (+)= A * B^T over `ksteps` slices of K = 16, bf16x3 with the six leading products
// (smallest first).  PA / PB: pieces present in A / B (1 for exact constants such as ones).
// Called by a whole CONVERGED warp; `leader` (umma::elect_one()) predicates the instructions.
// Completion is observed through umma::mma_commit by the same lane.
template <bool A_MN, bool B_MN, int PA = 3, int PB = 3>
__device__ __forceinline__ void gemm(bool leader, uint32_t d_tmem, const Mat& A, const Mat& B, int M, int N,
                                     int ksteps, uint32_t accumulate) {
  const uint32_t idesc = umma::idesc_bf16(M, N, A_MN ? 1 : 0, B_MN ? 1 : 0);
  uint64_t ad[PA], bd[PB];
#pragma unroll
  for (int i = 0; i < PA; ++i) ad[i] = make_desc(desc_lo<A_MN>(A) + i * (A.piece_bytes >> 4), desc_hi<A_MN>(A));
#pragma unroll
  for (int j = 0; j < PB; ++j) bd[j] = make_desc(desc_lo<B_MN>(B) + j * (B.piece_bytes >> 4), desc_hi<B_MN>(B));
  const uint64_t aks = desc_kstep<A_MN>(A), bks = desc_kstep<B_MN>(B);
  uint32_t acc = accumulate;
#pragma unroll 1
  for (int ks = 0; ks < ksteps; ++ks) {
#pragma unroll
    for (int sum = 2; sum >= 0; --sum) {
#pragma unroll
      for (int i = 0; i <= sum; ++i) {
        const int j = sum - i;
        if (i < PA && j < PB) {
          if (leader) umma::mma_f16(d_tmem, ad[i], bd[j], idesc, acc);
          acc = 1u;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < PA; ++i) ad[i] += aks;  // start-address field only; never carries out of it
#pragma unroll
    for (int j = 0; j < PB; ++j) bd[j] += bks;
  }
}

// Same with the A operand in TMEM: piece i of the 64-wide operand occupies 32 packed columns at
// a_tmem + 32 * i, a K = 16 slice is 8 columns.
template <bool B_MN, int PA = 3, int PB = 3>
__device__ __forceinline__ void gemm_ts(bool leader, uint32_t d_tmem, uint32_t a_tmem, const Mat& B, int M, int N,
                                        int ksteps, uint32_t accumulate) {
  const uint32_t idesc = umma::idesc_bf16(M, N, 0, B_MN ? 1 : 0);
  uint64_t bd[PB];
#pragma unroll
  for (int j = 0; j < PB; ++j) bd[j] = make_desc(desc_lo<B_MN>(B) + j * (B.piece_bytes >> 4), desc_hi<B_MN>(B));
  const uint64_t bks = desc_kstep<B_MN>(B);
  uint32_t acc = accumulate;
  uint32_t at = a_tmem;
#pragma unroll 1
  for (int ks = 0; ks < ksteps; ++ks) {
#pragma unroll
    for (int sum = 2; sum >= 0; --sum) {
#pragma unroll
      for (int i = 0; i <= sum; ++i) {
        const int j = sum - i;
        if (i < PA && j < PB) {
          if (leader) umma::mma_f16_ts(d_tmem, at + 32u * i, bd[j], idesc, acc);
          acc = 1u;
        }
      }
    }
    at += 8u;
#pragma unroll
    for (int j = 0; j < PB; ++j) bd[j] += bks;
  }
}

// Stage a Linear's weight W[out][in] (fp32, row-major, global) as the canonical matrix with
// r = out, c = in (zero-padded to in_pad, a multiple of 8).
__device__ __forceinline__ void stage_weight(const float* __restrict__ W, int out, int in, int in_pad, const Mat& dst,
                                             int tid, int nthreads) {
  const int chunks = in_pad >> 3;
  for (int e = tid; e < out * chunks; e += nthreads) {
    const int o = e % out, ch = e / out;
    float v[8];
    const float* src = W + (size_t)o * in + ch * 8;
    if (ch * 8 + 8 <= in && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      // one 32-byte sector per lane in two requests: eight scalar loads per lane hit eight times as many sectors
      // at once as the (mostly carved-out) L1 can hold, and the staging is on every minibatch's critical path
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

// The default policy this path is specialised for (policy/actor_critic/mlp.py:64-78 defaults):
// encoder d->64->64->16, value head 16->64->64->1, action head 16->64->16->n_out.
__host__ inline bool is_default_policy(const LrxPolicyDesc& p) {
  if (p.obs_dim < 1 || p.obs_dim > 8 || p.n_out < 1 || p.n_out > 8) return false;
  const int want[3][4] = {{p.obs_dim, 64, 64, 16}, {16, 64, 64, 1}, {16, 64, 16, p.n_out}};
  const int relu[3][3] = {{1, 1, 0}, {1, 1, 0}, {1, 0, 0}};
  for (int c = 0; c < 3; ++c) {
    if (p.n_layers[c] != 3) return false;
    for (int i = 0; i < 4; ++i)
      if (p.dims[c][i] != want[c][i]) return false;
    for (int i = 0; i < 3; ++i)
      if ((p.relu[c][i] != 0) != (relu[c][i] != 0)) return false;
  }
  return true;
}

// Offsets into the block of fp32 parameters kept in shared memory next to the bf16 weights.
enum : int {
  F_B0 = 0, F_B1 = 64, F_B2 = 128, F_BV0 = 144, F_BV1 = 208, F_BV2 = 272, F_BA0 = 276, F_BA1 = 340,
  F_BMAP = 356, F_WV2 = 364, F_WMAP = 428,
  F_W0 = 556,  // encoder layer 0 weight [64][8] (inputs zero-padded): that layer runs on the FMA pipe
  F_COUNT = 1068, F_RESERVE = 1088
};
// Layer order in the flat (Equinox leaf order) parameter vector.
enum : int { L_ENC0 = 0, L_ENC1, L_ENC2, L_V0, L_V1, L_V2, L_A0, L_A1, L_MAP, L_COUNT };

struct LayerOffsets {
  int w[L_COUNT], b[L_COUNT];
  int log_std;  // -1 for Discrete
};
__host__ inline LayerOffsets layer_offsets(const LrxPolicyDesc& p) {
  LayerOffsets o{};
  int off = 0, l = 0;
  for (int c = 0; c < 3; ++c)
    for (int i = 0; i < 3; ++i) {
      const int k = p.dims[c][i], n = p.dims[c][i + 1];
      o.w[l] = off;
      o.b[l] = off + k * n;
      off += k * n + n;
      ++l;
    }
  o.log_std = p.is_box ? off : -1;
  return o;
}

// Copy biases, the scalar value-head output row and the action mapping to shared fp32.
__device__ __forceinline__ void stage_f32(const float* __restrict__ params, const LayerOffsets& lo, int n_out, int obs_dim,
                                          float* fp, int tid, int nthreads) {
  for (int i = tid; i < F_COUNT; i += nthreads) {
    float v = 0.f;
    if (i < F_B1) v = params[lo.b[L_ENC0] + i];
    else if (i < F_B2) v = params[lo.b[L_ENC1] + (i - F_B1)];
    else if (i < F_BV0) v = params[lo.b[L_ENC2] + (i - F_B2)];
    else if (i < F_BV1) v = params[lo.b[L_V0] + (i - F_BV0)];
    else if (i < F_BV2) v = params[lo.b[L_V1] + (i - F_BV1)];
    else if (i < F_BA0) v = (i == F_BV2) ? params[lo.b[L_V2]] : 0.f;
    else if (i < F_BA1) v = params[lo.b[L_A0] + (i - F_BA0)];
    else if (i < F_BMAP) v = params[lo.b[L_A1] + (i - F_BA1)];
    else if (i < F_WV2) v = (i - F_BMAP) < n_out ? params[lo.b[L_MAP] + (i - F_BMAP)] : 0.f;
    else if (i < F_WMAP) v = params[lo.w[L_V2] + (i - F_WV2)];
    else if (i < F_W0) v = (i - F_WMAP) < n_out * 16 ? params[lo.w[L_MAP] + (i - F_WMAP)] : 0.f;
    else {
      const int o = (i - F_W0) >> 3, k = (i - F_W0) & 7;
      v = k < obs_dim ? params[lo.w[L_ENC0] + o * obs_dim + k] : 0.f;
    }
    fp[i] = v;
  }
}

}  // namespace tc
