// Error reporting, descriptor validation and misc C-ABI entry points.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/lerax_b200_debug.h"
#include "lrx_common.cuh"

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void lrx_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void lrx_count_launch(int n) { g_launches += n; }

extern "C" const char* lrx_last_error(void) { return g_err; }
extern "C" int lrx_version(void) { return LRX_VERSION; }
extern "C" int64_t lrx_launch_count(int reset) {
  int64_t v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

// Kernel selection for default-width policies (include/lerax_b200_debug.h): 2 = newest tensor-core kernels (default),
// 1 = first-generation tensor-core update kernel, 0 = generic-width kernels everywhere.  Initial value from the
// environment (LRX_FUSED=0..4, LRX_DISABLE_FUSED=1).  A test hook: the tests run the implementations against each other.
static int g_fused = -1;
int lrx_fused_enabled() {
  if (g_fused < 0) {
    const char* d = getenv("LRX_DISABLE_FUSED");
    const char* e = getenv("LRX_FUSED");
    g_fused = (d && d[0] == '1') ? 0 : (e && e[0] >= '0' && e[0] <= '4') ? e[0] - '0' : 4;
  }
  return g_fused;
}
// Programmatic dependent launch between the update's kernels (LRX_PDL=0 switches it off: an A/B knob)
int lrx_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LRX_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v;
}
extern "C" void lrx_debug_set_fused(int level) { g_fused = level < 0 ? 0 : level > 4 ? 4 : level; }

int lrx_sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      n <= 0)
    return 148;
  return n < LRX_MAX_SMS ? n : LRX_MAX_SMS;
}

int lrx_env_state_dim(int kind) { return (kind == LRX_ENV_CARTPOLE || kind == LRX_ENV_ACROBOT) ? 4 : 2; }
int lrx_env_obs_dim(int kind) {
  return kind == LRX_ENV_CARTPOLE ? 4 : kind == LRX_ENV_PENDULUM ? 3 : kind == LRX_ENV_ACROBOT ? 6 : 2;
}
bool lrx_env_is_box(int kind) { return kind == LRX_ENV_PENDULUM || kind == LRX_ENV_CONTINUOUS_MOUNTAINCAR; }
int lrx_env_num_outputs(int kind) {  // policy head width: number of discrete actions, 1 for a scalar Box
  return kind == LRX_ENV_CARTPOLE ? 2 : (kind == LRX_ENV_ACROBOT || kind == LRX_ENV_MOUNTAINCAR) ? 3 : 1;
}

int lrx_validate_env(const LrxEnvDesc* env) {
  LRX_REQUIRE(env != nullptr, LRX_ERR_INVALID_ARG, "NULL LrxEnvDesc");
  LRX_REQUIRE(env->kind >= LRX_ENV_CARTPOLE && env->kind <= LRX_ENV_CONTINUOUS_MOUNTAINCAR, LRX_ERR_UNSUPPORTED,
              "unsupported env kind %d (supported: CartPole, Pendulum, Acrobot, MountainCar, ContinuousMountainCar)",
              env->kind);
  LRX_REQUIRE(env->dt > 0.0f, LRX_ERR_INVALID_ARG, "env dt must be positive");
  LRX_REQUIRE(env->max_episode_steps >= 0, LRX_ERR_INVALID_ARG, "max_episode_steps must be >= 0");
  return LRX_OK;
}

int lrx_validate_policy(const LrxPolicyDesc* pol) {
  LRX_REQUIRE(pol != nullptr, LRX_ERR_INVALID_ARG, "NULL LrxPolicyDesc");
  LRX_REQUIRE(pol->obs_dim >= 1 && pol->n_out >= 1 && pol->n_out <= LRX_MAX_NOUT, LRX_ERR_UNSUPPORTED,
              "obs_dim/n_out out of range (n_out <= %d)", LRX_MAX_NOUT);
  LRX_REQUIRE(!pol->is_box || pol->n_out == 1, LRX_ERR_UNSUPPORTED,
              "only scalar Box actions are supported (n_out must be 1)");
  for (int c = 0; c < 3; ++c) {
    LRX_REQUIRE(pol->n_layers[c] >= 1 && pol->n_layers[c] <= LRX_MAX_LAYERS, LRX_ERR_UNSUPPORTED,
                "chain %d: n_layers %d out of range [1, %d]", c, pol->n_layers[c], LRX_MAX_LAYERS);
    for (int i = 0; i <= pol->n_layers[c]; ++i)
      LRX_REQUIRE(pol->dims[c][i] >= 1 && pol->dims[c][i] <= 4096, LRX_ERR_UNSUPPORTED,
                  "chain %d: layer width %d out of range", c, pol->dims[c][i]);
  }
  const int feat = pol->dims[0][pol->n_layers[0]];
  LRX_REQUIRE(pol->dims[0][0] == pol->obs_dim, LRX_ERR_INVALID_ARG, "encoder input width != obs_dim");
  LRX_REQUIRE(pol->dims[1][0] == feat && pol->dims[2][0] == feat, LRX_ERR_INVALID_ARG,
              "value/action head input width != feature_size");
  LRX_REQUIRE(pol->dims[1][pol->n_layers[1]] == 1, LRX_ERR_INVALID_ARG, "value head must end in a scalar");
  LRX_REQUIRE(pol->dims[2][pol->n_layers[2]] == pol->n_out, LRX_ERR_INVALID_ARG, "action head must end in n_out");
  int np = 0;
  lrx_policy_layers(*pol, nullptr, nullptr, &np);
  LRX_REQUIRE(pol->n_params == np, LRX_ERR_INVALID_ARG, "n_params %d != %d implied by the layer widths",
              pol->n_params, np);
  return LRX_OK;
}

extern "C" int lrx_policy_num_params(const LrxPolicyDesc* pol) {
  if (!pol) {
    lrx_set_error("NULL LrxPolicyDesc");
    return LRX_ERR_INVALID_ARG;
  }
  int np = 0;
  lrx_policy_layers(*pol, nullptr, nullptr, &np);
  return np;
}
