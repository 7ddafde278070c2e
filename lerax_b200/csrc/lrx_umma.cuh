// tcgen05 / TMEM / mbarrier building blocks for the sm_100a tensor-core paths (inline PTX).
//
// All matrices that feed the tensor core live in shared memory in ONE canonical
// no-swizzle core-matrix layout ("chunk-major"): element (r, c) of an fp32 matrix with R
// rows is at byte
//         (c / 4) * (R * 16)  +  r * 16  +  (c % 4) * 4 .
// A 16-byte chunk holds 4 consecutive c of one r; the R chunks of one c-group are
// contiguous (so a warp whose lanes own consecutive r writes 512 contiguous bytes:
// conflict-free STS.128).  The same bytes serve as
//   * a K-major operand  (MN index = r, K index = c):  LBO = R*16, SBO = 128
//   * an MN-major operand (MN index = c, K index = r): LBO = 128,  SBO = R*16
// which is what lets the weight-gradient GEMMs (contraction over the batch rows) read the
// activations the forward pass wrote without any transpose.
//
// fp32 fidelity: kind::tf32 reads 32-bit containers and ignores the low 13 mantissa bits,
// so every product is issued three times ("3xTF32"): a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo
// with x_hi = x with the low 13 mantissa bits cleared (exactly representable in tf32, so
// the result does not depend on how the hardware would round) and x_lo = x - x_hi, each in
// its own buffer.  Accumulation is fp32 in TMEM.  Relative error ~2^-21 per product.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// Warp-uniform helpers: the MMA issue path runs with the whole warp converged (so descriptor
// arithmetic can live in uniform registers) and only the instruction itself is predicated on
// one elected lane.
__device__ __forceinline__ int warp_idx_sync() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xFFFFFFFF;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(count), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
// generic-proxy smem writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------- TMEM
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// One full warp allocates `ncols` (power of two >= 32) columns; base address -> *slot (smem).
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMEM address = (lane << 16) | column.  A warp may only touch lanes [32*(warp%4), +32).
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, uint32_t lane, uint32_t col) {
  return base + (lane << 16) + col;
}

// tcgen05.ld 32x32b: thread i of the warp receives lane (base_lane + i), registers = columns.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// tcgen05.ld 16x256b.x2: 16 TMEM lanes x 16 columns spread over all 32 threads (measured with
// tools/tmem_shape_probe.py): thread t holds lanes r0 = t/4 and r1 = r0 + 8, columns c = 2*(t%4):
//   v[0],v[1] = (r0, c), (r0, c+1)   v[2],v[3] = (r1, c), (r1, c+1)
//   v[4],v[5] = (r0, c+8), (r0, c+9) v[6],v[7] = (r1, c+8), (r1, c+9)
__device__ __forceinline__ void tmem_ld_frag16(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, SWIZZLE_NONE, sm_100 version field = 1.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type = 0) {
  uint64_t d = (uint64_t)(layout_type & 7u) << 61;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// Instruction descriptor for kind::tf32, fp32 accumulate.  a_mn / b_mn: 1 = MN-major operand.
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Instruction descriptor for kind::f16 with bf16 operands, fp32 accumulate.
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// A operand from TMEM (bf16 pairs packed in 32-bit columns; for M = 64 row m sits in TMEM lane
// m%16 + 32*(m/16), word c holds K elements 2c, 2c+1 -- tools/tmem_operand_probe.py), B from shared memory.
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// tcgen05.st 16x128b.x1: thread t writes r0 -> (lane t/4, column t%4), r1 -> (lane t/4 + 8, column t%4)
// (tools/tmem_st_shape_probe.py): the store twin of the 16x256b fragment for packed bf16 pairs.
__device__ __forceinline__ void tmem_st_frag(uint32_t taddr, uint32_t r0, uint32_t r1) {
  asm volatile("tcgen05.st.sync.aligned.16x128b.x1.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(r0), "r"(r1) : "memory");
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on `bar` once every previously issued tcgen05.mma of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// An operand in the canonical layout: `addr` = shared address of element (0,0), R = rows of
// the stored matrix (16-byte units per c-group), mn = 1 when the GEMM's M/N index runs along
// c (the stored matrix is used transposed).
struct Operand {
  uint32_t addr_hi, addr_lo;  // raw fp32 copy and its low-part copy
  uint32_t R;
  int mn;
};
__device__ __forceinline__ uint64_t operand_desc(uint32_t addr, uint32_t R, int mn, int kstep) {
  // K-major: one MMA consumes K = 8 = two c-groups; MN-major: K = 8 rows of one c-group set.
  return mn ? smem_desc(addr + (uint32_t)kstep * 128u, 128u, R * 16u)
            : smem_desc(addr + (uint32_t)kstep * 2u * R * 16u, R * 16u, 128u);
}

// D[tmem] (+)= A * B^T over K (multiple of 8) with the 3xTF32 split.  One thread issues.
// accumulate: 0 -> the first MMA overwrites D.
__device__ __forceinline__ void gemm_tf32x3(uint32_t d_tmem, const Operand& A, const Operand& B, int M, int N,
                                            int K, uint32_t accumulate) {
  const uint32_t idesc = idesc_tf32(M, N, A.mn, B.mn);
  uint32_t acc = accumulate;
  for (int ks = 0; ks < K / 8; ++ks) {
    const uint64_t ah = operand_desc(A.addr_hi, A.R, A.mn, ks), al = operand_desc(A.addr_lo, A.R, A.mn, ks);
    const uint64_t bh = operand_desc(B.addr_hi, B.R, B.mn, ks), bl = operand_desc(B.addr_lo, B.R, B.mn, ks);
    mma_tf32(d_tmem, al, bh, idesc, acc);  // small terms first
    mma_tf32(d_tmem, ah, bl, idesc, 1u);
    mma_tf32(d_tmem, ah, bh, idesc, 1u);
    acc = 1u;
  }
}
__device__ __forceinline__ void gemm_tf32x1(uint32_t d_tmem, const Operand& A, const Operand& B, int M, int N,
                                            int K, uint32_t accumulate) {
  const uint32_t idesc = idesc_tf32(M, N, A.mn, B.mn);
  uint32_t acc = accumulate;
  for (int ks = 0; ks < K / 8; ++ks) {
    mma_tf32(d_tmem, operand_desc(A.addr_hi, A.R, A.mn, ks), operand_desc(B.addr_hi, B.R, B.mn, ks), idesc, acc);
    acc = 1u;
  }
}

// x = hi + lo with hi = the 19 bits kind::tf32 keeps, lo = x - hi exact in fp32.
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ float tf32_lo(float x) { return x - tf32_hi(x); }

// Byte offset of element (r, c) in the canonical layout of an R-row matrix.
__host__ __device__ constexpr uint32_t canon_off(uint32_t R, uint32_t r, uint32_t c) {
  return (c >> 2) * (R * 16u) + r * 16u + (c & 3u) * 4u;
}

}  // namespace umma
