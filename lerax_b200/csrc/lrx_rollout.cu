// Rollout path: env reset / step / observe, batched policy forward, and the fused
// persistent rollout kernel (generic-width version: one lane per env, activations in
// shared memory columns, weights through the read-only path).
//
// Reference: algorithm/ppo.py:199-316 (PPO.step / collect_rollout), :360-368 (vmap over envs).
#include "lrx_env.cuh"
#include "lrx_policy.cuh"

constexpr uint32_t LRX_STREAM_INIT = 2;

// ============================================================== small batched env kernels
template <int KIND>
__global__ void env_reset_kernel(LrxEnvDesc envd, LrxEnvState st, int N, LrxRng rng,
                                 const float* __restrict__ uniform) {
  using E = Env<KIND>;
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  E env;
  env.load(envd);
  float u[E::SD], y[E::SD];
  if (uniform) {
#pragma unroll
    for (int i = 0; i < E::SD; ++i) u[i] = uniform[(size_t)n * E::SD + i];
  } else {
    Philox4 w = lrx_noise_words(rng, n, 0, LRX_STREAM_INIT);
    uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < E::SD; ++i) u[i] = lrx_u32_to_unit(ww[i]);
  }
  env.initial(u, y);
#pragma unroll
  for (int i = 0; i < E::SD; ++i) st.y[(size_t)i * N + n] = y[i];
  st.t[n] = 0.0f;
  st.step_count[n] = 0;
}

template <int KIND>
__global__ void env_step_kernel(LrxEnvDesc envd, LrxEnvState st, const void* __restrict__ action,
                                float* reward, uint8_t* terminal, uint8_t* truncated,
                                float* obs_next, int N) {
  using E = Env<KIND>;
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  E env;
  env.load(envd);
  float y[E::SD], y1[E::SD], t1;
#pragma unroll
  for (int i = 0; i < E::SD; ++i) y[i] = st.y[(size_t)i * N + n];
  EnvAction a;
  a.i = EnvDims<KIND>::BOX ? 0 : ((const int32_t*)action)[n];
  a.f = EnvDims<KIND>::BOX ? ((const float*)action)[n] : 0.f;
  env.transition(y, st.t[n], a, y1, t1);
  int sc = st.step_count[n] + 1;
  if (reward) reward[n] = env.reward(a, y, y1);
  if (terminal) terminal[n] = env.terminal(y1) ? 1 : 0;
  if (truncated) truncated[n] = env.truncate(sc) ? 1 : 0;
  if (obs_next) {
    float o[E::OD];
    env.observe(y1, o);
#pragma unroll
    for (int i = 0; i < E::OD; ++i) obs_next[(size_t)n * E::OD + i] = o[i];
  }
#pragma unroll
  for (int i = 0; i < E::SD; ++i) st.y[(size_t)i * N + n] = y1[i];
  st.t[n] = t1;
  st.step_count[n] = sc;
}

template <int KIND>
__global__ void env_observe_kernel(LrxEnvDesc envd, LrxEnvState st, float* obs, uint8_t* terminal, int N) {
  using E = Env<KIND>;
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  E env;
  env.load(envd);
  float y[E::SD], o[E::OD];
#pragma unroll
  for (int i = 0; i < E::SD; ++i) y[i] = st.y[(size_t)i * N + n];
  if (obs) {
    env.observe(y, o);
#pragma unroll
    for (int i = 0; i < E::OD; ++i) obs[(size_t)n * E::OD + i] = o[i];
  }
  if (terminal) terminal[n] = env.terminal(y) ? 1 : 0;
}

// ================================================================ batched policy forward
__global__ void policy_forward_kernel(const __grid_constant__ PolicyTable tab,
                                      const float* __restrict__ params,
                                      const float* __restrict__ obs, float* value, float* head,
                                      int B, int obs_dim, int n_out) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per_warp = (2 * tab.max_width + tab.feat) * 32;
  float* bufA = smem + (size_t)warp * per_warp;
  float* bufB = bufA + tab.max_width * 32;
  float* feat = bufB + tab.max_width * 32;
  const int warps_per_block = blockDim.x >> 5;
  for (int base = (blockIdx.x * warps_per_block + warp) * 32; base < B;
       base += gridDim.x * warps_per_block * 32) {
    int m = base + lane;
    int ml = m < B ? m : B - 1;
    for (int k = 0; k < obs_dim; ++k) bufA[k * 32 + lane] = obs[(size_t)ml * obs_dim + k];
    float* f = lane_chain(params, tab, 0, bufA, bufA, bufB, lane);
    for (int k = 0; k < tab.feat; ++k) feat[k * 32 + lane] = f[k * 32 + lane];
    if (value) {
      float* v = lane_chain(params, tab, 1, feat, bufA, bufB, lane);
      if (m < B) value[m] = v[lane];
    }
    if (head) {
      float* h = lane_chain(params, tab, 2, feat, bufA, bufB, lane);
      if (m < B)
        for (int j = 0; j < n_out; ++j) head[(size_t)m * n_out + j] = h[j * 32 + lane];
    }
  }
}

// ===================================================================== rollout (generic)
// One lane = one env for all T steps; env state stays in registers, the lane's activation
// column stays in shared memory; nothing but the emitted rows touches HBM.
template <int KIND>
__global__ void __launch_bounds__(256)
rollout_generic_kernel(const __grid_constant__ LrxEnvDesc envd, const __grid_constant__ PolicyTable tab,
                       const float* __restrict__ params, LrxEnvState st, LrxRng rng,
                       LrxReplay rp, LrxRolloutOut out, int N, int T, float gamma, int n_out) {
  using E = Env<KIND>;
  constexpr int SD = E::SD, OD = E::OD;
  constexpr bool BOX = EnvDims<KIND>::BOX;
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per_warp = (2 * tab.max_width + tab.feat) * 32;
  float* bufA = smem + (size_t)warp * per_warp;
  float* bufB = bufA + tab.max_width * 32;
  float* feat = bufB + tab.max_width * 32;

  const int n = (blockIdx.x * (blockDim.x >> 5) + warp) * 32 + lane;
  const bool live = n < N;
  const int nl = live ? n : N - 1;

  E env;
  env.load(envd);
  float y[SD], t;
  int sc;
#pragma unroll
  for (int i = 0; i < SD; ++i) y[i] = st.y[(size_t)i * N + nl];
  t = st.t[nl];
  sc = st.step_count[nl];

  float log_std = 0.f, scale = 1.f, log_scale = 0.f;
  if (BOX) {
    log_std = __ldg(params + tab.log_std_off);
    scale = expf(log_std);     // actor.py:68-72: Normal(loc, exp(log_std))
    log_scale = logf(scale);   // distreqx takes log(scale)
  }

  for (int s = 0; s < T; ++s) {
    const size_t row = (size_t)s * N + nl;
    if (out.y_trace && live) {
#pragma unroll
      for (int i = 0; i < SD; ++i) out.y_trace[((size_t)s * SD + i) * N + n] = y[i];
    }
    if (out.t_trace && live) out.t_trace[row] = t;

    // ---- observation of the PRE-transition state, policy.action_and_value (mlp.py:133-152)
    float o[OD];
    env.observe(y, o);
#pragma unroll
    for (int k = 0; k < OD; ++k) bufA[k * 32 + lane] = o[k];
    float* f = lane_chain(params, tab, 0, bufA, bufA, bufB, lane);
    for (int k = 0; k < tab.feat; ++k) feat[k * 32 + lane] = f[k * 32 + lane];
    float value = lane_chain(params, tab, 1, feat, bufA, bufB, lane)[lane];
    float* hd = lane_chain(params, tab, 2, feat, bufA, bufB, lane);

    EnvAction a;
    a.i = 0; a.f = 0.f;
    float logp;
    if (BOX) {
      float loc = hd[lane];
      float eps;
      if (rp.action_noise) {
        eps = rp.action_noise[row];
      } else {
        Philox4 w = lrx_noise_words(rng, n, s, LRX_STREAM_ACTION);
        float u1 = lrx_u32_to_unit(w.x), u2 = lrx_u32_to_unit(w.y);
        eps = sqrtf(-2.0f * logf(u1)) * cosf(6.28318530717958647692f * u2);
      }
      float a_raw = loc + scale * eps;
      logp = normal_log_prob(loc, scale, log_scale, a_raw);  // log-prob of the UNCLIPPED sample
      a.f = lrx_clipf(a_raw, env.action_low(), env.action_high());  // ppo.py:228-235 (Box bounds of the env)
    } else {
      float z[LRX_MAX_NOUT];
#pragma unroll
      for (int j = 0; j < LRX_MAX_NOUT; ++j) z[j] = (j < n_out) ? hd[j * 32 + lane] : 0.f;
      log_softmax_n<LRX_MAX_NOUT>(z, n_out);
      Philox4 w{0u, 0u, 0u, 0u};
      if (!rp.action_noise) w = lrx_noise_words(rng, n, s, LRX_STREAM_ACTION);
      uint32_t ww[4] = {w.x, w.y, w.z, w.w};
      float best = -INFINITY;
      int bi = 0;
#pragma unroll
      for (int j = 0; j < LRX_MAX_NOUT; ++j) {
        if (j < n_out) {
          float g;
          if (rp.action_noise) {
            g = rp.action_noise[row * n_out + j];
          } else {
            if (j >= 4 && (j & 3) == 0) {
              Philox4 w2 = philox4x32_10(rng.env_offset + n, rng.step_offset + s, LRX_STREAM_ACTION,
                                         (uint32_t)(j >> 2), (uint32_t)rng.seed, (uint32_t)(rng.seed >> 32));
              ww[0] = w2.x; ww[1] = w2.y; ww[2] = w2.z; ww[3] = w2.w;
            }
            g = -logf(-logf(lrx_u32_to_unit(ww[j & 3])));
          }
          float v = z[j] + g;
          if (v > best) { best = v; bi = j; }  // first maximum wins (jnp.argmax)
        }
      }
      a.i = bi;
      logp = 0.f;
#pragma unroll
      for (int j = 0; j < LRX_MAX_NOUT; ++j) if (j == bi) logp = z[j];
    }

    // ---- env.transition / reward / terminal / truncate (ppo.py:237-246)
    float y1[SD], t1;
    env.transition(y, t, a, y1, t1);
    int sc1 = sc + 1;
    float r = env.reward(a, y, y1);
    bool term = env.terminal(y1);
    bool trunc = env.truncate(sc1);
    bool done = term || trunc;
    if (trunc) {
      // ppo.py:248-259: r += gamma * V(obs(next_state)) with the PRE-reset next state
      float o1[OD];
      env.observe(y1, o1);
#pragma unroll
      for (int k = 0; k < OD; ++k) bufA[k * 32 + lane] = o1[k];
      float* f1 = lane_chain(params, tab, 0, bufA, bufA, bufB, lane);
      for (int k = 0; k < tab.feat; ++k) feat[k * 32 + lane] = f1[k * 32 + lane];
      float vn = lane_chain(params, tab, 1, feat, bufA, bufB, lane)[lane];
      r = r + gamma * vn;
    }
    if (out.y_next_trace && live) {
#pragma unroll
      for (int i = 0; i < SD; ++i) out.y_next_trace[((size_t)s * SD + i) * N + n] = y1[i];
    }

    // ---- emit the buffer row (ppo.py:276-288)
    if (live) {
#pragma unroll
      for (int k = 0; k < OD; ++k) out.obs[row * OD + k] = o[k];
      if (BOX) ((float*)out.act)[row] = a.f; else ((int32_t*)out.act)[row] = a.i;
      out.rew[row] = r;
      out.done[row] = done ? 1 : 0;
      out.logp[row] = logp;
      out.val[row] = value;
    }

    // ---- reset-on-done (ppo.py:261-263): env.initial(fresh noise), t = 0, step_count = 0
    if (done) {
      float u[SD];
      if (rp.reset_uniform) {
#pragma unroll
        for (int i = 0; i < SD; ++i) u[i] = rp.reset_uniform[row * SD + i];
      } else {
        Philox4 w = lrx_noise_words(rng, n, s, LRX_STREAM_RESET);
        uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int i = 0; i < SD; ++i) u[i] = lrx_u32_to_unit(ww[i]);
      }
      env.initial(u, y);
      t = 0.f;
      sc = 0;
    } else {
#pragma unroll
      for (int i = 0; i < SD; ++i) y[i] = y1[i];
      t = t1;
      sc = sc1;
    }
  }

  // ---- last_value = V(obs(final state)) (ppo.py:311-312)
  {
    float o[OD];
    env.observe(y, o);
#pragma unroll
    for (int k = 0; k < OD; ++k) bufA[k * 32 + lane] = o[k];
    float* f = lane_chain(params, tab, 0, bufA, bufA, bufB, lane);
    for (int k = 0; k < tab.feat; ++k) feat[k * 32 + lane] = f[k * 32 + lane];
    float lv = lane_chain(params, tab, 1, feat, bufA, bufB, lane)[lane];
    if (live) {
      out.last_value[n] = lv;
#pragma unroll
      for (int i = 0; i < SD; ++i) st.y[(size_t)i * N + n] = y[i];
      st.t[n] = t;
      st.step_count[n] = sc;
    }
  }
}

// ================================================================================ host
static int warps_for_smem(const PolicyTable& tab, int max_warps, size_t budget) {
  size_t per_warp = (size_t)(2 * tab.max_width + tab.feat) * 32 * sizeof(float);
  int w = (int)(budget / per_warp);
  if (w < 1) w = 1;
  if (w > max_warps) w = max_warps;
  return w;
}

#define DISPATCH_ENV(kind, ...)                                    \
  switch (kind) {                                                  \
    case LRX_ENV_CARTPOLE: { constexpr int K_ = LRX_ENV_CARTPOLE; __VA_ARGS__; } break; \
    case LRX_ENV_PENDULUM: { constexpr int K_ = LRX_ENV_PENDULUM; __VA_ARGS__; } break; \
    case LRX_ENV_ACROBOT:  { constexpr int K_ = LRX_ENV_ACROBOT;  __VA_ARGS__; } break; \
    case LRX_ENV_MOUNTAINCAR: { constexpr int K_ = LRX_ENV_MOUNTAINCAR; __VA_ARGS__; } break; \
    case LRX_ENV_CONTINUOUS_MOUNTAINCAR: { constexpr int K_ = LRX_ENV_CONTINUOUS_MOUNTAINCAR; __VA_ARGS__; } break; \
    default: lrx_set_error("unknown env kind %d", kind); return LRX_ERR_UNSUPPORTED; \
  }

extern "C" int lrx_env_reset(const LrxEnvDesc* env, LrxEnvState state, int32_t N, LrxRng rng,
                             const float* uniform, lrx_stream_t stream) {
  int rc = lrx_validate_env(env);
  if (rc) return rc;
  LRX_REQUIRE(N >= 0 && state.y && state.t && state.step_count, LRX_ERR_INVALID_ARG,
              "lrx_env_reset: bad N or NULL state");
  if (N == 0) return LRX_OK;
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH_ENV(env->kind, (env_reset_kernel<K_><<<(N + 255) / 256, 256, 0, s>>>(*env, state, N, rng, uniform)));
  LRX_LAUNCH_CHECK();
  return LRX_OK;
}

extern "C" int lrx_env_step(const LrxEnvDesc* env, LrxEnvState state, const void* action,
                            float* reward, uint8_t* terminal, uint8_t* truncated, float* obs_next,
                            int32_t N, lrx_stream_t stream) {
  int rc = lrx_validate_env(env);
  if (rc) return rc;
  LRX_REQUIRE(N >= 0 && state.y && state.t && state.step_count && action, LRX_ERR_INVALID_ARG,
              "lrx_env_step: bad N or NULL state/action");
  if (N == 0) return LRX_OK;
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH_ENV(env->kind, (env_step_kernel<K_><<<(N + 255) / 256, 256, 0, s>>>(
                              *env, state, action, reward, terminal, truncated, obs_next, N)));
  LRX_LAUNCH_CHECK();
  return LRX_OK;
}

extern "C" int lrx_env_observe(const LrxEnvDesc* env, LrxEnvState state, float* obs, uint8_t* terminal,
                               int32_t N, lrx_stream_t stream) {
  int rc = lrx_validate_env(env);
  if (rc) return rc;
  LRX_REQUIRE(N >= 0 && state.y && (obs || terminal), LRX_ERR_INVALID_ARG, "lrx_env_observe: bad N or NULL pointer");
  if (N == 0) return LRX_OK;
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH_ENV(env->kind, (env_observe_kernel<K_><<<(N + 255) / 256, 256, 0, s>>>(*env, state, obs, terminal, N)));
  LRX_LAUNCH_CHECK();
  return LRX_OK;
}

extern "C" int lrx_policy_forward(const LrxPolicyDesc* pol, const float* params, const float* obs,
                                  float* value, float* head, int32_t B, lrx_stream_t stream) {
  int rc = lrx_validate_policy(pol);
  if (rc) return rc;
  LRX_REQUIRE(B >= 0 && params && obs, LRX_ERR_INVALID_ARG, "lrx_policy_forward: bad B or NULL pointer");
  if (B == 0) return LRX_OK;
  PolicyTable tab = lrx_make_table(*pol);
  int warps = warps_for_smem(tab, 8, 96 * 1024);
  size_t smem = (size_t)warps * (2 * tab.max_width + tab.feat) * 32 * sizeof(float);
  LRX_REQUIRE(smem <= 227 * 1024, LRX_ERR_UNSUPPORTED, "policy too wide for shared memory (%zu B)", smem);
  LRX_CUDA_CHECK(cudaFuncSetAttribute(policy_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int blocks = (B + warps * 32 - 1) / (warps * 32);
  if (blocks > lrx_sm_count() * 8) blocks = lrx_sm_count() * 8;
  policy_forward_kernel<<<blocks, warps * 32, smem, (cudaStream_t)stream>>>(tab, params, obs, value, head, B,
                                                                           pol->obs_dim, pol->n_out);
  LRX_LAUNCH_CHECK();
  return LRX_OK;
}

int lrx_rollout_fast(const LrxEnvDesc* env, const LrxPolicyDesc* pol, const float* params,
                     LrxEnvState state, LrxRng rng, const LrxReplay* replay, LrxRolloutOut out,
                     int32_t N, int32_t T, float gamma, cudaStream_t stream, bool* handled);

int lrx_rollout_stepwise(const LrxEnvDesc* env, const LrxPolicyDesc* pol, const float* params, LrxEnvState state,
                         LrxRng rng, const LrxReplay* replay, LrxRolloutOut out, int32_t N, int32_t T, float gamma,
                         cudaStream_t s, bool* handled);

extern "C" int lrx_rollout(const LrxEnvDesc* env, const LrxPolicyDesc* pol, const float* params,
                           LrxEnvState state, LrxRng rng, const LrxReplay* replay, LrxRolloutOut out,
                           int32_t N, int32_t T, float gamma, lrx_stream_t stream) {
  int rc = lrx_validate_env(env);
  if (rc) return rc;
  rc = lrx_validate_policy(pol);
  if (rc) return rc;
  LRX_REQUIRE(N >= 0 && T >= 0, LRX_ERR_INVALID_ARG, "lrx_rollout: negative N or T");
  LRX_REQUIRE(params && state.y && state.t && state.step_count, LRX_ERR_INVALID_ARG,
              "lrx_rollout: NULL params/state");
  LRX_REQUIRE(out.last_value && (T == 0 || (out.obs && out.act && out.rew && out.done && out.logp && out.val)),
              LRX_ERR_INVALID_ARG, "lrx_rollout: NULL output plane");
  LRX_REQUIRE(pol->obs_dim == lrx_env_obs_dim(env->kind), LRX_ERR_INVALID_ARG,
              "policy obs_dim %d does not match env obs dim %d", pol->obs_dim, lrx_env_obs_dim(env->kind));
  const bool box = lrx_env_is_box(env->kind);
  LRX_REQUIRE((pol->is_box != 0) == box, LRX_ERR_INVALID_ARG, "policy action type does not match env");
  const int na = lrx_env_num_outputs(env->kind);
  LRX_REQUIRE(pol->n_out == na, LRX_ERR_INVALID_ARG, "policy n_out %d does not match env (%d)", pol->n_out, na);
  if (N == 0) return LRX_OK;
  cudaStream_t s = (cudaStream_t)stream;
  LrxReplay rp{nullptr, nullptr};
  if (replay) rp = *replay;

  bool handled = false;
  rc = lrx_rollout_fast(env, pol, params, state, rng, &rp, out, N, T, gamma, s, &handled);
  if (rc) return rc;
  if (handled) return LRX_OK;
  rc = lrx_rollout_stepwise(env, pol, params, state, rng, &rp, out, N, T, gamma, s, &handled);  // wide policies, many envs
  if (rc) return rc;
  if (handled) return LRX_OK;

  PolicyTable tab = lrx_make_table(*pol);
  int warps = warps_for_smem(tab, 8, 96 * 1024);
  // spread the envs over all SMs before stacking warps into a block
  while (warps > 1 && (N + warps * 32 - 1) / (warps * 32) < lrx_sm_count()) warps >>= 1;
  size_t smem = (size_t)warps * (2 * tab.max_width + tab.feat) * 32 * sizeof(float);
  LRX_REQUIRE(smem <= 227 * 1024, LRX_ERR_UNSUPPORTED, "policy too wide for shared memory (%zu B)", smem);
  int blocks = (N + warps * 32 - 1) / (warps * 32);
  DISPATCH_ENV(env->kind, {
    LRX_CUDA_CHECK(cudaFuncSetAttribute(rollout_generic_kernel<K_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rollout_generic_kernel<K_><<<blocks, warps * 32, smem, s>>>(*env, tab, params, state, rng, rp, out, N, T, gamma, pol->n_out);
  });
  LRX_LAUNCH_CHECK();
  return LRX_OK;
}
