// Fused per-warp PPO gradient kernel for the default policy widths (tried before the
// generic layer-by-layer path).
#include "lrx_common.cuh"
#include "lrx_policy.cuh"

size_t lrx_fused_workspace_bytes(const LrxPolicyDesc* pol, int64_t B) { return 0; }

int lrx_ppo_grad_fused(const LrxPolicyDesc* pol, const float* params, const LrxBatchView* buf,
                       const int32_t* idx, int64_t B, int64_t B_global, const double* adv_sums,
                       const LrxPpoHyper* h, float* gradbuf, void* workspace, size_t workspace_bytes,
                       cudaStream_t stream, bool* handled) {
  *handled = false;
  return LRX_OK;
}
