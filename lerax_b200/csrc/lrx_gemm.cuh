// C[M x N] (+)= op(A)[M x K] * op(B)[K x N] for the layer-by-layer (generic-width) PPO update: one argument block
// for the FMA kernel (sgemm_kernel, lrx_update.cu) and the tcgen05 kernel (lrx_gemm_tc.cu).
#pragma once

#include <cuda_runtime.h>

struct GemmArgs {
  const float* A; int lda;
  const float* B; int ldb;
  float* C; int ldc;
  int M, N, K;
  const float* bias;   // + bias[i]
  int relu;
  const float* mask; int ldm;  // *= (mask[i][j] > 0)
  int accumulate;      // C += result
  int ones_row;        // additionally write C[M][j] = 1
  int k_chunk;         // K range per blockIdx.z (split-K)
  long long c_split_stride;  // C/C2 offset per blockIdx.z
  float* C2; int n_main;     // column j == n_main is redirected to C2[i]
};

// tcgen05 path; *handled = false when the shape stays on the FMA kernel.
int lrx_launch_gemm_tc(const GemmArgs& g, bool at, bool bt, int splits, cudaStream_t s, bool* handled);
