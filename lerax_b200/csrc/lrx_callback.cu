// Built-in `on_step` callback states recomputed from the (reward, done) streams of a rollout
// (SURVEY 8f-2).  In the reference `LoggingCallback.on_step` runs inside the per-step scan
// (callback/logging/callback.py:477-480) and advances `LoggingCallbackStepState.next`
// (:139-176) per env; that state is a pure function of the reward / done sequence, so one
// forward scan over the [T][N] buffer the fused rollout kernel wrote reproduces it exactly.
// One thread per env (coalesced across envs at every step); 5 bytes per env-step.
#include "lrx_common.cuh"

__global__ void episode_stats_kernel(const float* __restrict__ rew, const uint8_t* __restrict__ done, LrxEpisodeStats st,
                                     int N, int T, float alpha, int layout) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  long long step = st.step[n];
  float ep_ret = st.episode_return[n], avg_ret = st.average_return[n], avg_len = st.average_length[n];
  int ep_len = st.episode_length[n];
  bool ep_done = st.episode_done[n] != 0;
  const float one_minus_alpha = __fsub_rn(1.0f, alpha);
  for (int t = 0; t < T; ++t) {
    const size_t i = layout == LRX_LAYOUT_TN ? (size_t)t * N + n : (size_t)n * T + t;
    const float r = rew[i];
    const bool d = done[i] != 0;
    // callback.py:152-155: the previous step's done flag resets the running episode sums
    ep_ret = __fadd_rn(__fmul_rn(ep_ret, ep_done ? 0.0f : 1.0f), r);
    ep_len = (ep_done ? 0 : ep_len) + 1;
    if (d) {  // callback.py:157-166 (no fused multiply-add: the oracle rounds the products)
      avg_ret = __fadd_rn(__fmul_rn(alpha, ep_ret), __fmul_rn(one_minus_alpha, avg_ret));
      avg_len = __fadd_rn(__fmul_rn(alpha, (float)ep_len), __fmul_rn(one_minus_alpha, avg_len));
    }
    ep_done = d;
    ++step;
  }
  st.step[n] = step;
  st.episode_return[n] = ep_ret;
  st.episode_length[n] = ep_len;
  st.episode_done[n] = ep_done ? 1 : 0;
  st.average_return[n] = avg_ret;
  st.average_length[n] = avg_len;
}

extern "C" int lrx_episode_stats(const float* rew, const uint8_t* done, const LrxEpisodeStats* st, int32_t N, int32_t T,
                                 float alpha, int32_t layout, lrx_stream_t stream) {
  LRX_REQUIRE(N >= 0 && T >= 0, LRX_ERR_INVALID_ARG, "lrx_episode_stats: negative size");
  LRX_REQUIRE(layout == LRX_LAYOUT_TN || layout == LRX_LAYOUT_NT, LRX_ERR_INVALID_ARG, "lrx_episode_stats: bad layout");
  if (N == 0 || T == 0) return LRX_OK;
  LRX_REQUIRE(rew && done && st && st->step && st->episode_return && st->episode_length && st->episode_done &&
                  st->average_return && st->average_length,
              LRX_ERR_INVALID_ARG, "lrx_episode_stats: NULL pointer");
  episode_stats_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rew, done, *st, N, T, alpha, layout);
  LRX_LAUNCH_CHECK();
  return LRX_OK;
}
