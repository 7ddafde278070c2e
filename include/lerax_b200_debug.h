/* lerax_b200 -- TEST HOOKS.  Not part of the drop-in boundary (include/lerax_b200.h).
 *
 * Two groups:
 *  (1) process-wide toggles exported by liblerax_b200.so itself, used by the tests to run two implementations of one
 *      entry point against each other and by tools/ to sweep launch shapes.  They are the only process-wide mutable
 *      state of the library; production callers never touch them.
 *  (2) self-tests and probes of the tcgen05 / TMEM building blocks, built into the separate
 *      lerax_b200/_C/liblerax_b200_test.so (csrc/lrx_umma_selftest.cu), which links against the product library.
 */
#ifndef LERAX_B200_DEBUG_H
#define LERAX_B200_DEBUG_H

#include "lerax_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ----------------------------------------------------------------- (1) hooks in liblerax_b200.so
 * Kernel selection: 2 (default) = default-width policies run the newest tensor-core kernels; 1 = the first-generation
 * tensor-core update kernel (bf16x3, M = 64 tiles); 0 = everything runs the generic-width kernels.  Initial value
 * from the environment variable LRX_FUSED (0 / 1 / 2); LRX_DISABLE_FUSED=1 is the same as LRX_FUSED=0. */
void lrx_debug_set_fused(int level);
/* GAE launch shape for lrx_gae(LRX_LAYOUT_TN): knob = EPL * 10000 + L * 100 + C (envs per lane, steps per thread,
 * warps per block); 0 = automatic.  tools/gae_time.py sweeps it. */
void lrx_debug_set_gae_chunk(int knob);
/* Cycle profile of ppo_grad_tc128_kernel: with a device buffer of 64 int64 set, CTA 0 accumulates the cycles between
 * consecutive step marks of row-owner thread 0 (slots 1..16) and of the issuer warp (slots 33..48); NULL switches it off. */
void lrx_debug_set_tc128_prof(long long* device_buffer_64);
/* ppo_grad_pp64_kernel: raw clock64 stamps of CTA 0 / half 0, 8 per issue step, first `records` steps of a launch. */
void lrx_debug_set_pp64_prof(long long* device_buffer, int records);
/* device-to-device copy of n floats to a raw device address (an exchange slot of lrx_p2p_*). */
/* The generic-width update's GEMM by itself: C[M x N] (+)= op(A)[M x K] op(B)[K x N] (+ bias[i], ReLU, ReLU mask,
 * split-K over gridDim.z into C + z * c_split_stride); use_tc = 0 forces the fp32 FMA kernel. */
int lrx_debug_gemm(const float* A, int32_t lda, const float* B, int32_t ldb, float* C, int32_t ldc, int32_t M, int32_t N,
                   int32_t K, int32_t at, int32_t bt, const float* bias, int32_t relu, const float* mask, int32_t ldm,
                   int32_t accumulate, int32_t splits, int32_t k_chunk, int64_t c_split_stride, int32_t use_tc,
                   lrx_stream_t stream);
int lrx_debug_copy_f32(float* dst, const float* src, int32_t n, lrx_stream_t stream);

/* ----------------------------------------------------------------- (2) liblerax_b200_test.so */
/* ------------------------------------------------------------------ diagnostics
 * Self-test of the tensor-core building blocks the fused kernels are made of (tcgen05.mma
 * kind::tf32 with the 3xTF32 split, operands in the canonical no-swizzle shared-memory
 * layout, accumulator in TMEM): D = A[M][K] * B[N][K]^T on ONE SM.  M in {64,128};
 * N multiple of 8 (16 when M = 128); K multiple of 8.  cfg_host (HOST pointer, 12 ints):
 * {mn, s_mn, s_k, lbo, sbo, kstep} for A then B -- mn = 1 stages that operand MN-major (the
 * transposed reading the weight-gradient GEMMs use), s_* are the storage strides between
 * core-matrix groups and lbo/sbo/kstep the descriptor fields, all in bytes.  split = 1 -> 3xTF32.
 * dump receives the raw accumulator [128 TMEM lanes][N].  Not part of the reference path. */
int lrx_debug_umma_gemm(const float* A, const float* B, float* dump, int32_t M, int32_t N, int32_t K,
                        const int32_t* cfg_host, int32_t split, lrx_stream_t stream);

/* Same for the bf16x3 fp32 emulation (kind::f16, bf16 pieces x = p0+p1+p2, nprod = 1, 3 or 6
 * products kept; K multiple of 16; 2-byte elements, 8 per 16-byte chunk). */
int lrx_debug_umma_gemm_bf16(const float* A, const float* B, float* dump, int32_t M, int32_t N, int32_t K,
                             const int32_t* cfg_host, int32_t nprod, lrx_stream_t stream);

/* Timing probe for the same MMA sequence: cycles[0] = issue, cycles[1] = issue + completion. */
int lrx_debug_umma_time_bf16(const float* A, const float* B, float* dump, int32_t M, int32_t N, int32_t K,
                             const int32_t* cfg_host, int32_t nprod, int32_t reps, long long* cycles,
                             lrx_stream_t stream);

/* Dispatch-rate probe: 256 back-to-back MMAs with constant descriptors over `dsplit` accumulators. */
int lrx_debug_umma_rate(int32_t M, int32_t N, int32_t dsplit, long long* cycles, lrx_stream_t stream);
/* Latency of tcgen05.ld from one TMEM region while n_mma MMAs accumulate into another (cycles[0..7] per load,
 * cycles[8] until the MMA batch completes).  dsplit bits 8 / 9 of lrx_debug_umma_rate select MN-major A / B. */
int lrx_debug_tmem_ld_under_mma(int32_t n_mma, long long* cycles, lrx_stream_t stream);

/* Decode of the tcgen05.ld 16x256b.x2 fragment shape: dump [128 threads][8 registers]. */
int lrx_debug_tmem_shape_probe(float* dump, lrx_stream_t stream);


/* Probes of the A-operand-in-TMEM layout (mode 0) and of an M = 64 accumulator at TMEM lane 16 (mode 1):
 * dump [128 threads][16 columns] of the accumulator. */
int lrx_debug_tmem_operand_probe(float* dump, int32_t mode, int32_t variant, int32_t M, lrx_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LERAX_B200_DEBUG_H */
