// Rollout for NON-default policy widths (BASELINE.json configs[2]: 256-wide MLPs), batched over the envs step by step.
//
// Reference: algorithm/ppo.py:199-316 (PPO.step scanned over T under the vmap of :360-368) -- the same event order as
// rollout_generic_kernel (lrx_rollout.cu).  That kernel gives every env one lane for all T steps: fine while the
// weights fit the read-only cache, but a 256-wide policy is 148 k MACs per env-step on a single lane and 16 384 envs
// fill 512 warps of a 9 472-warp machine (1.6 TFLOP/s measured).  Here one step of ALL envs is one batch:
//     observe -> [d][N] ; every Linear of the three chains = one GEMM  Out[out][N] = act(W[out][in] In[in][N] + b)
//     -> one kernel per step samples the action, takes the env transition, reward / terminal / truncate, (TimeLimit:
//     a second value forward on the pre-reset next state for the bootstrap of ppo.py:248-259), resets and stores the
//     buffer row.
// The GEMMs are the exact-fp32 FMA kernel of the update path (the rollout's contract is 1e-5 on values and log-probs,
// tighter than the two-piece tensor-core emulation), N = envs is the parallel axis.  Activations live in a
// caller-allocated workspace (LrxRolloutOut.workspace, lrx_rollout_workspace_bytes).
#include "lrx_env.cuh"
#include "lrx_gemm.cuh"
#include "lrx_policy.cuh"

int lrx_launch_gemm_fma(const GemmArgs& g, bool at, bool bt, int splits, cudaStream_t s);  // lrx_update.cu

namespace {

struct StepWs {
  int ld;
  float* obsT;                          // [d][ld]
  float* act[3][LRX_MAX_LAYERS];        // [out][ld] per Linear
  float* y1;                            // [SD][ld] post-transition, PRE-reset state (TimeLimit path)
  float* t1;                            // [ld]
  float* rew;                           // [ld]
  int32_t* sc1;                         // [ld]
  uint8_t* flags;                       // [ld] bit 0 terminal, bit 1 truncated
  size_t total;
};

size_t align_up_s(size_t x, size_t a) { return (x + a - 1) / a * a; }

StepWs carve_steps(const LrxPolicyDesc& pol, int N, void* base) {
  StepWs w{};
  w.ld = (N + 3) / 4 * 4;
  char* b = (char*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) { void* p = b ? b + off : nullptr; off += align_up_s(bytes, 256); return p; };
  w.obsT = (float*)take((size_t)pol.obs_dim * w.ld * 4);
  for (int c = 0; c < 3; ++c)
    for (int i = 0; i < pol.n_layers[c]; ++i) w.act[c][i] = (float*)take((size_t)pol.dims[c][i + 1] * w.ld * 4);
  w.y1 = (float*)take((size_t)4 * w.ld * 4);
  w.t1 = (float*)take((size_t)w.ld * 4);
  w.rew = (float*)take((size_t)w.ld * 4);
  w.sc1 = (int32_t*)take((size_t)w.ld * 4);
  w.flags = (uint8_t*)take((size_t)w.ld);
  w.total = off;
  return w;
}

// env.observation of the carried state -> obsT [d][ld] (and the buffer row / traces of step s when `row_obs` is given)
template <int KIND>
__global__ void __launch_bounds__(256)
observe_kernel(LrxEnvDesc envd, const float* __restrict__ y_in, int y_ld, const float* __restrict__ t_in, int N,
               float* __restrict__ obsT, int ld, float* row_obs, float* y_trace, float* t_trace) {
  using E = Env<KIND>;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  E env;
  env.load(envd);
  float y[E::SD], o[E::OD];
#pragma unroll
  for (int i = 0; i < E::SD; ++i) y[i] = y_in[(size_t)i * y_ld + n];
  env.observe(y, o);
#pragma unroll
  for (int k = 0; k < E::OD; ++k) obsT[(size_t)k * ld + n] = o[k];
  if (row_obs) {
#pragma unroll
    for (int k = 0; k < E::OD; ++k) row_obs[(size_t)n * E::OD + k] = o[k];
  }
  if (y_trace) {
#pragma unroll
    for (int i = 0; i < E::SD; ++i) y_trace[(size_t)i * N + n] = y[i];
  }
  if (t_trace) t_trace[n] = t_in[n];
}

struct ActArgs {
  LrxEnvDesc envd;
  LrxEnvState st;
  LrxRng rng;
  LrxReplay rp;
  LrxRolloutOut out;
  const float* head;   // [n_out][ld]
  const float* value;  // [ld]
  const float* params;
  int log_std_off, n_out, N, ld, s;
  float gamma;
  int two_phase;       // TimeLimit: stop after the transition, finish_kernel completes the step
  float* y1; float* t1; float* rew; int32_t* sc1; uint8_t* flags; float* obsT;
};

template <int KIND>
__device__ __forceinline__ void finish_row(const Env<KIND>& env, const ActArgs& a, int n, size_t row, float r, bool done,
                                           const float (&y1)[EnvDims<KIND>::SD], float t1, int sc1) {
  constexpr int SD = EnvDims<KIND>::SD;
  a.out.rew[row] = r;
  a.out.done[row] = done ? 1 : 0;
  float y[SD], t;
  int sc;
  if (done) {  // ppo.py:261-263: env.initial(fresh noise), t = 0, step_count = 0
    float u[SD];
    if (a.rp.reset_uniform) {
#pragma unroll
      for (int i = 0; i < SD; ++i) u[i] = a.rp.reset_uniform[row * SD + i];
    } else {
      Philox4 w = lrx_noise_words(a.rng, n, a.s, LRX_STREAM_RESET);
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
#pragma unroll
  for (int i = 0; i < SD; ++i) a.st.y[(size_t)i * a.N + n] = y[i];
  a.st.t[n] = t;
  a.st.step_count[n] = sc;
}

// policy.action_and_value's sampling half (mlp.py:133-152, distreqx Categorical / Normal), ppo.py:228-246: clip, transition,
// reward, terminal, truncate; then either the rest of the step (no TimeLimit) or the hand-over to finish_kernel.
template <int KIND>
__global__ void __launch_bounds__(256) act_kernel(const __grid_constant__ ActArgs a) {
  using E = Env<KIND>;
  constexpr int SD = E::SD, OD = E::OD;
  constexpr bool BOX = EnvDims<KIND>::BOX;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.N) return;
  E env;
  env.load(a.envd);
  const size_t row = (size_t)a.s * a.N + n;
  float y[SD];
#pragma unroll
  for (int i = 0; i < SD; ++i) y[i] = a.st.y[(size_t)i * a.N + n];
  const float t = a.st.t[n];
  const int sc = a.st.step_count[n];
  const float value = a.value[n];

  EnvAction act;
  act.i = 0; act.f = 0.f;
  float logp;
  if (BOX) {
    const float log_std = __ldg(a.params + a.log_std_off);
    const float scale = expf(log_std), log_scale = logf(scale);
    const float loc = a.head[n];
    float eps;
    if (a.rp.action_noise) {
      eps = a.rp.action_noise[row];
    } else {
      Philox4 w = lrx_noise_words(a.rng, n, a.s, LRX_STREAM_ACTION);
      float u1 = lrx_u32_to_unit(w.x), u2 = lrx_u32_to_unit(w.y);
      eps = sqrtf(-2.0f * logf(u1)) * cosf(6.28318530717958647692f * u2);
    }
    const float a_raw = loc + scale * eps;
    logp = normal_log_prob(loc, scale, log_scale, a_raw);  // log-prob of the UNCLIPPED sample
    act.f = lrx_clipf(a_raw, env.action_low(), env.action_high());
  } else {
    float z[LRX_MAX_NOUT];
#pragma unroll
    for (int j = 0; j < LRX_MAX_NOUT; ++j) z[j] = (j < a.n_out) ? a.head[(size_t)j * a.ld + n] : 0.f;
    log_softmax_n<LRX_MAX_NOUT>(z, a.n_out);
    Philox4 w{0u, 0u, 0u, 0u};
    if (!a.rp.action_noise) w = lrx_noise_words(a.rng, n, a.s, LRX_STREAM_ACTION);
    uint32_t ww[4] = {w.x, w.y, w.z, w.w};
    float best = -INFINITY;
    int bi = 0;
#pragma unroll
    for (int j = 0; j < LRX_MAX_NOUT; ++j) {
      if (j < a.n_out) {
        float g;
        if (a.rp.action_noise) {
          g = a.rp.action_noise[row * a.n_out + j];
        } else {
          if (j >= 4 && (j & 3) == 0) {
            Philox4 w2 = philox4x32_10(a.rng.env_offset + n, a.rng.step_offset + a.s, LRX_STREAM_ACTION, (uint32_t)(j >> 2),
                                       (uint32_t)a.rng.seed, (uint32_t)(a.rng.seed >> 32));
            ww[0] = w2.x; ww[1] = w2.y; ww[2] = w2.z; ww[3] = w2.w;
          }
          g = -logf(-logf(lrx_u32_to_unit(ww[j & 3])));
        }
        const float v = z[j] + g;
        if (v > best) { best = v; bi = j; }  // first maximum wins (jnp.argmax)
      }
    }
    act.i = bi;
    logp = 0.f;
#pragma unroll
    for (int j = 0; j < LRX_MAX_NOUT; ++j) if (j == bi) logp = z[j];
  }

  float y1[SD], t1;
  env.transition(y, t, act, y1, t1);
  const int sc1 = sc + 1;
  const float r = env.reward(act, y, y1);
  const bool term = env.terminal(y1), trunc = env.truncate(sc1);
  if (a.out.y_next_trace) {
#pragma unroll
    for (int i = 0; i < SD; ++i) a.out.y_next_trace[((size_t)a.s * SD + i) * a.N + n] = y1[i];
  }
  if (BOX) ((float*)a.out.act)[row] = act.f; else ((int32_t*)a.out.act)[row] = act.i;
  a.out.logp[row] = logp;
  a.out.val[row] = value;
  if (!a.two_phase) {
    finish_row<KIND>(env, a, n, row, r, term || trunc, y1, t1, sc1);
    return;
  }
  // TimeLimit: the bootstrap needs V(obs(next state)) BEFORE the reset (ppo.py:248-259): hand the next state over
  float o1[OD];
  env.observe(y1, o1);
#pragma unroll
  for (int k = 0; k < OD; ++k) a.obsT[(size_t)k * a.ld + n] = o1[k];
#pragma unroll
  for (int i = 0; i < SD; ++i) a.y1[(size_t)i * a.ld + n] = y1[i];
  a.t1[n] = t1;
  a.sc1[n] = sc1;
  a.rew[n] = r;
  a.flags[n] = (term ? 1 : 0) | (trunc ? 2 : 0);
}

template <int KIND>
__global__ void __launch_bounds__(256) finish_kernel(const __grid_constant__ ActArgs a) {
  using E = Env<KIND>;
  constexpr int SD = E::SD;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.N) return;
  E env;
  env.load(a.envd);
  const size_t row = (size_t)a.s * a.N + n;
  float y1[SD];
#pragma unroll
  for (int i = 0; i < SD; ++i) y1[i] = a.y1[(size_t)i * a.ld + n];
  const uint8_t f = a.flags[n];
  float r = a.rew[n];
  if (f & 2) r = r + a.gamma * a.value[n];  // a.value: V of the pre-reset next observation
  finish_row<KIND>(env, a, n, row, r, f != 0, y1, a.t1[n], a.sc1[n]);
}

__global__ void copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// one chain of Linears: in [k][ld] -> act[c][last] ; returns the output buffer
int run_chain(const LrxPolicyDesc& pol, const PolicyTable& tab, const float* params, const StepWs& w, int c, const float* in,
              int N, cudaStream_t s, const float** out) {
  for (int i = 0; i < tab.count[c]; ++i) {
    const LrxLayer& L = tab.L[tab.first[c] + i];
    GemmArgs g{};
    g.A = params + L.w_off; g.lda = L.in;
    g.B = in; g.ldb = w.ld;
    g.C = w.act[c][i]; g.ldc = w.ld;
    g.M = L.out; g.N = N; g.K = L.in;
    g.bias = params + L.b_off; g.relu = L.relu;
    g.k_chunk = L.in;
    int rc = lrx_launch_gemm_fma(g, false, false, 1, s);
    if (rc) return rc;
    in = w.act[c][i];
  }
  *out = in;
  return LRX_OK;
}

}  // namespace

extern "C" size_t lrx_rollout_workspace_bytes(const LrxPolicyDesc* pol, int32_t N) {
  if (!pol || lrx_validate_policy(pol) != LRX_OK || N <= 0) return 0;
  return carve_steps(*pol, N, nullptr).total;
}

#define STEP_DISPATCH(kind, ...)                                    \
  switch (kind) {                                                  \
    case LRX_ENV_CARTPOLE: { constexpr int K_ = LRX_ENV_CARTPOLE; __VA_ARGS__; } break; \
    case LRX_ENV_PENDULUM: { constexpr int K_ = LRX_ENV_PENDULUM; __VA_ARGS__; } break; \
    case LRX_ENV_ACROBOT:  { constexpr int K_ = LRX_ENV_ACROBOT;  __VA_ARGS__; } break; \
    case LRX_ENV_MOUNTAINCAR: { constexpr int K_ = LRX_ENV_MOUNTAINCAR; __VA_ARGS__; } break; \
    case LRX_ENV_CONTINUOUS_MOUNTAINCAR: { constexpr int K_ = LRX_ENV_CONTINUOUS_MOUNTAINCAR; __VA_ARGS__; } break; \
    default: lrx_set_error("unknown env kind %d", kind); return LRX_ERR_UNSUPPORTED; \
  }

int lrx_rollout_stepwise(const LrxEnvDesc* env, const LrxPolicyDesc* pol, const float* params, LrxEnvState state,
                         LrxRng rng, const LrxReplay* replay, LrxRolloutOut out, int32_t N, int32_t T, float gamma,
                         cudaStream_t s, bool* handled) {
  *handled = false;
  if (!out.workspace || lrx_fused_enabled() < 1 || N < 512) return LRX_OK;  // small batches: one lane per env is fine
  const StepWs w = carve_steps(*pol, N, out.workspace);
  LRX_REQUIRE(out.workspace_bytes >= w.total, LRX_ERR_INVALID_ARG, "lrx_rollout: workspace too small (%zu < %zu)",
              (size_t)out.workspace_bytes, w.total);
  const PolicyTable tab = lrx_make_table(*pol);
  const int blocks = (N + 255) / 256;
  const int SDk = lrx_env_state_dim(env->kind), ODk = lrx_env_obs_dim(env->kind);
  ActArgs a{};
  a.envd = *env; a.st = state; a.rng = rng; a.rp = *replay; a.out = out; a.params = params;
  a.log_std_off = tab.log_std_off; a.n_out = pol->n_out; a.N = N; a.ld = w.ld; a.gamma = gamma;
  a.two_phase = env->max_episode_steps > 0;
  a.y1 = w.y1; a.t1 = w.t1; a.rew = w.rew; a.sc1 = w.sc1; a.flags = w.flags; a.obsT = w.obsT;
  const float *feat, *val, *head;
  int rc;
  for (int t = 0; t < T; ++t) {
    float* row_obs = out.obs + (size_t)t * N * ODk;
    float* ytr = out.y_trace ? out.y_trace + (size_t)t * SDk * N : nullptr;
    float* ttr = out.t_trace ? out.t_trace + (size_t)t * N : nullptr;
    STEP_DISPATCH(env->kind, (observe_kernel<K_><<<blocks, 256, 0, s>>>(*env, state.y, N, state.t, N, w.obsT, w.ld, row_obs, ytr, ttr)));
    LRX_LAUNCH_CHECK();
    if ((rc = run_chain(*pol, tab, params, w, 0, w.obsT, N, s, &feat))) return rc;
    if ((rc = run_chain(*pol, tab, params, w, 1, feat, N, s, &val))) return rc;
    if ((rc = run_chain(*pol, tab, params, w, 2, feat, N, s, &head))) return rc;
    a.s = t; a.head = head; a.value = val;
    STEP_DISPATCH(env->kind, (act_kernel<K_><<<blocks, 256, 0, s>>>(a)));
    LRX_LAUNCH_CHECK();
    if (a.two_phase) {
      if ((rc = run_chain(*pol, tab, params, w, 0, w.obsT, N, s, &feat))) return rc;
      if ((rc = run_chain(*pol, tab, params, w, 1, feat, N, s, &val))) return rc;
      a.value = val;
      STEP_DISPATCH(env->kind, (finish_kernel<K_><<<blocks, 256, 0, s>>>(a)));
      LRX_LAUNCH_CHECK();
    }
  }
  // last_value = V(obs(final state)) (ppo.py:311-312)
  STEP_DISPATCH(env->kind, (observe_kernel<K_><<<blocks, 256, 0, s>>>(*env, state.y, N, state.t, N, w.obsT, w.ld, nullptr, nullptr, nullptr)));
  LRX_LAUNCH_CHECK();
  if ((rc = run_chain(*pol, tab, params, w, 0, w.obsT, N, s, &feat))) return rc;
  if ((rc = run_chain(*pol, tab, params, w, 1, feat, N, s, &val))) return rc;
  copy_kernel<<<blocks, 256, 0, s>>>(val, out.last_value, N);
  LRX_LAUNCH_CHECK();
  *handled = true;
  return LRX_OK;
}
