// Register-tiled rollout for the default policy widths (tried before the generic kernel).
#include "lrx_env.cuh"
#include "lrx_policy.cuh"

int lrx_rollout_fast(const LrxEnvDesc* env, const LrxPolicyDesc* pol, const float* params,
                     LrxEnvState state, LrxRng rng, const LrxReplay* replay, LrxRolloutOut out,
                     int32_t N, int32_t T, float gamma, cudaStream_t stream, bool* handled) {
  *handled = false;
  return LRX_OK;
}
