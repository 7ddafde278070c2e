// Classic-control environments as device functors: dynamics, one Tsit5 step, clip,
// observation, reward, terminal, initial.  One thread owns one env; everything lives in
// registers.  Each function cites the reference lines it re-implements (paths relative to
// /root/reference/src/lerax); operation order follows the reference expressions.
#pragma once

#include "lrx_common.cuh"

template <int KIND> struct EnvDims;
template <> struct EnvDims<LRX_ENV_CARTPOLE> { static constexpr int SD = 4, OD = 4, NA = 2; static constexpr bool BOX = false; };
template <> struct EnvDims<LRX_ENV_PENDULUM> { static constexpr int SD = 2, OD = 3, NA = 1; static constexpr bool BOX = true; };
template <> struct EnvDims<LRX_ENV_ACROBOT>  { static constexpr int SD = 4, OD = 6, NA = 3; static constexpr bool BOX = false; };
template <> struct EnvDims<LRX_ENV_MOUNTAINCAR> { static constexpr int SD = 2, OD = 2, NA = 3; static constexpr bool BOX = false; };
template <> struct EnvDims<LRX_ENV_CONTINUOUS_MOUNTAINCAR> { static constexpr int SD = 2, OD = 2, NA = 1; static constexpr bool BOX = true; };

// An action is carried as both views; discrete envs read .i, Box envs read .f.
struct EnvAction {
  int i;
  float f;
};

// Tsitouras 5(4) tableau (diffrax 0.7.2 Tsit5; oracle/tsit5.py), rounded to fp32 as JAX
// does when the float64 tableau meets fp32 state.
#define TS_A21 0.161f
#define TS_A31 (-0.008480655492356989f)
#define TS_A32 0.335480655492357f
#define TS_A41 2.8971530571054935f
#define TS_A42 (-6.359448489975075f)
#define TS_A43 4.3622954328695815f
#define TS_A51 5.325864828439257f
#define TS_A52 (-11.748883564062828f)
#define TS_A53 7.4955393428898365f
#define TS_A54 (-0.09249506636175525f)
#define TS_A61 5.86145544294642f
#define TS_A62 (-12.92096931784711f)
#define TS_A63 8.159367898576159f
#define TS_A64 (-0.071584973281401f)
#define TS_A65 (-0.028269050394068383f)
#define TS_B1 0.09646076681806523f
#define TS_B2 0.01f
#define TS_B3 0.4798896504144996f
#define TS_B4 1.379008574103742f
#define TS_B5 (-3.290069515436081f)
#define TS_B6 2.324710524099774f

// jnp.remainder(x, m) for m > 0: C fmod shifted into [0, m).
__device__ __forceinline__ float lrx_modf_pos(float x, float m) {
  float r = fmodf(x, m);
  return (r < 0.0f) ? r + m : r;
}

__device__ __forceinline__ float lrx_wrap_angle(float th) {
  const float PI_F = 3.14159265358979323846f, TWO_PI_F = 6.28318530717958647692f;
  return lrx_modf_pos(th + PI_F, TWO_PI_F) - PI_F;
}

template <int KIND>
struct Env {
  using D = EnvDims<KIND>;
  static constexpr int SD = D::SD, OD = D::OD;
  float c[LRX_ENV_MAX_PARAMS];  // physical params (+ derived constants at the tail)
  float dt;
  int max_steps;

  __device__ __forceinline__ void load(const LrxEnvDesc& e) {
#pragma unroll
    for (int i = 0; i < LRX_ENV_MAX_PARAMS; ++i) c[i] = e.p[i];
    dt = e.dt;
    max_steps = e.max_episode_steps;
    if (KIND == LRX_ENV_CARTPOLE) {
      c[7] = c[2] + c[1];  // total_mass = pole_mass + cart_mass      cartpole.py:127
      c[8] = c[2] * c[3];  // polemass_length = pole_mass * length    cartpole.py:129
    }
  }

  // ---- Box action bounds (space of the env; ppo.py:228-235 clips the sampled action to them)
  __device__ __forceinline__ float action_low() const {
    return KIND == LRX_ENV_PENDULUM ? -c[1] : KIND == LRX_ENV_CONTINUOUS_MOUNTAINCAR ? c[0] : 0.f;
  }
  __device__ __forceinline__ float action_high() const {
    return KIND == LRX_ENV_PENDULUM ? c[1] : KIND == LRX_ENV_CONTINUOUS_MOUNTAINCAR ? c[1] : 0.f;
  }

  // ---- dynamics: cartpole.py:162-177, pendulum.py:89-98, acrobot.py:167-232
  __device__ __forceinline__ void dynamics(const float (&y)[SD], const EnvAction& a,
                                           float (&dy)[SD]) const {
    if constexpr (KIND == LRX_ENV_CARTPOLE) {
      const float g = c[0], m_p = c[2], len = c[3], fmag = c[4], total = c[7], pml = c[8];
      float x_dot = y[1], theta = y[2], theta_dot = y[3];
      float force = (float)(a.i * 2 - 1) * fmag;
      float s, co;
      sincosf(theta, &s, &co);
      float temp = (force + pml * (theta_dot * theta_dot) * s) / total;
      float theta_dd = (g * s - co * temp) / (len * (1.3333333333333333f - m_p * (co * co) / total));
      float x_dd = temp - pml * theta_dd * co / total;
      dy[0] = x_dot; dy[1] = x_dd; dy[2] = theta_dot; dy[3] = theta_dd;
    } else if constexpr (KIND == LRX_ENV_PENDULUM) {
      const float mt = c[1], g = c[2], m = c[3], l = c[4];
      float u = lrx_clipf(a.f, -mt, mt);
      float theta_dd = 3.0f * g / (2.0f * l) * sinf(y[0]) + 3.0f / (m * (l * l)) * u;
      dy[0] = y[1]; dy[1] = theta_dd;
    } else if constexpr (KIND == LRX_ENV_MOUNTAINCAR) {
      // mountain_car.py:152-158
      const float u = (float)(a.i - 1) * c[5];
      const float x_dd = u - c[6] * cosf(3.0f * y[0]);
      dy[0] = y[1]; dy[1] = x_dd;
    } else if constexpr (KIND == LRX_ENV_CONTINUOUS_MOUNTAINCAR) {
      // continuous_mountain_car.py:138-144
      const float act = lrx_clipf(a.f, c[0], c[1]);
      const float x_dd = c[7] * act - 0.0025f * cosf(3.0f * y[0]);
      dy[0] = y[1]; dy[1] = x_dd;
    } else {
      const float g = c[0], l1 = c[1], m1 = c[3], m2 = c[4], lc1 = c[5], lc2 = c[6], moi = c[7];
      const float HALF_PI = 1.57079632679489661923f;
      float th1 = y[0], th2 = y[1], th1d = y[2], th2d = y[3];
      float tq = c[10 + a.i];
      float s2, c2;
      sincosf(th2, &s2, &c2);
      float d1 = m1 * (lc1 * lc1) + m2 * ((l1 * l1) + (lc2 * lc2) + 2.0f * l1 * lc2 * c2) + moi + moi;
      float d2 = m2 * ((lc2 * lc2) + l1 * lc2 * c2) + moi;
      float phi2 = m2 * lc2 * g * cosf(th1 + th2 - HALF_PI);
      float phi1 = -m2 * l1 * lc2 * (th2d * th2d) * s2 - 2.0f * m2 * l1 * lc2 * th1d * th2d * s2 +
                   (m1 * lc1 + m2 * l1) * g * cosf(th1 - HALF_PI) + phi2;
      float th2dd = (tq + d2 / d1 * phi1 - m2 * l1 * lc2 * (th1d * th1d) * s2 - phi2) /
                    (m2 * (lc2 * lc2) + moi - (d2 * d2) / d1);
      float th1dd = -(d2 * th2dd + phi1) / d1;
      dy[0] = th1d; dy[1] = th2d; dy[2] = th1dd; dy[3] = th2dd;
    }
  }

  // ---- env.clip: cartpole.py:179-180, pendulum.py:100-104, acrobot.py:234-241
  __device__ __forceinline__ void clip(float (&y)[SD]) const {
    if constexpr (KIND == LRX_ENV_PENDULUM) {
      y[0] = lrx_wrap_angle(y[0]);
      y[1] = lrx_clipf(y[1], -c[0], c[0]);
    } else if constexpr (KIND == LRX_ENV_ACROBOT) {
      y[0] = lrx_wrap_angle(y[0]);
      y[1] = lrx_wrap_angle(y[1]);
      y[2] = lrx_clipf(y[2], -c[8], c[8]);
      y[3] = lrx_clipf(y[3], -c[9], c[9]);
    } else if constexpr (KIND == LRX_ENV_MOUNTAINCAR) {
      // mountain_car.py:160-165: the car stops dead at the left wall
      float v = lrx_clipf(y[1], -c[2], c[2]);
      const float x = lrx_clipf(y[0], c[0], c[1]);
      v = v * (((x != c[0]) || (v > 0.0f)) ? 1.0f : 0.0f);
      y[0] = x; y[1] = v;
    } else if constexpr (KIND == LRX_ENV_CONTINUOUS_MOUNTAINCAR) {
      // continuous_mountain_car.py:146-150
      y[1] = lrx_clipf(y[1], -c[4], c[4]);
      y[0] = lrx_clipf(y[0], c[2], c[3]);
    }
  }

  // ---- transition: base_classic_control.py:49-72.  One Tsit5 step with h = (t+dt)-t
  // (diffrax ODETerm.contr(t0,t1) = t1 - t0 in fp32), k_i = f(y_i)*h, then clip, t += dt.
  __device__ __forceinline__ void transition(const float (&y0)[SD], float t, const EnvAction& a,
                                             float (&y1)[SD], float& t1) const {
    t1 = t + dt;
    const float h = t1 - t;
    float k1[SD], k2[SD], k3[SD], k4[SD], k5[SD], k6[SD], ys[SD], f[SD];
    dynamics(y0, a, f);
#pragma unroll
    for (int i = 0; i < SD; ++i) { k1[i] = f[i] * h; ys[i] = y0[i] + TS_A21 * k1[i]; }
    dynamics(ys, a, f);
#pragma unroll
    for (int i = 0; i < SD; ++i) { k2[i] = f[i] * h; ys[i] = y0[i] + (TS_A31 * k1[i] + TS_A32 * k2[i]); }
    dynamics(ys, a, f);
#pragma unroll
    for (int i = 0; i < SD; ++i) {
      k3[i] = f[i] * h;
      ys[i] = y0[i] + (TS_A41 * k1[i] + TS_A42 * k2[i] + TS_A43 * k3[i]);
    }
    dynamics(ys, a, f);
#pragma unroll
    for (int i = 0; i < SD; ++i) {
      k4[i] = f[i] * h;
      ys[i] = y0[i] + (TS_A51 * k1[i] + TS_A52 * k2[i] + TS_A53 * k3[i] + TS_A54 * k4[i]);
    }
    dynamics(ys, a, f);
#pragma unroll
    for (int i = 0; i < SD; ++i) {
      k5[i] = f[i] * h;
      ys[i] = y0[i] + (TS_A61 * k1[i] + TS_A62 * k2[i] + TS_A63 * k3[i] + TS_A64 * k4[i] + TS_A65 * k5[i]);
    }
    dynamics(ys, a, f);
#pragma unroll
    for (int i = 0; i < SD; ++i) {
      k6[i] = f[i] * h;
      y1[i] = y0[i] + (TS_B1 * k1[i] + TS_B2 * k2[i] + TS_B3 * k3[i] + TS_B4 * k4[i] + TS_B5 * k5[i] + TS_B6 * k6[i]);
    }
    clip(y1);
  }

  // ---- observation: cartpole.py:182-185, pendulum.py:106-110, acrobot.py:243-256
  __device__ __forceinline__ void observe(const float (&y)[SD], float (&o)[OD]) const {
    if constexpr (KIND == LRX_ENV_CARTPOLE) {
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = y[i];
    } else if constexpr (KIND == LRX_ENV_PENDULUM) {
      float s, co;
      sincosf(y[0], &s, &co);
      o[0] = co; o[1] = s; o[2] = y[1];
    } else if constexpr (KIND == LRX_ENV_MOUNTAINCAR || KIND == LRX_ENV_CONTINUOUS_MOUNTAINCAR) {
      o[0] = y[0]; o[1] = y[1];  // mountain_car.py:167-170, continuous_mountain_car.py:152-155
    } else {
      float s1, c1, s2, c2;
      sincosf(y[0], &s1, &c1);
      sincosf(y[1], &s2, &c2);
      o[0] = c1; o[1] = s1; o[2] = c2; o[3] = s2; o[4] = y[2]; o[5] = y[3];
    }
  }

  // ---- reward(state, action, next_state): cartpole.py:187-195, pendulum.py:112-123, acrobot.py:258-270,
  // mountain_car.py:172-180, continuous_mountain_car.py:157-168 (the latter tests the PRE-transition state)
  __device__ __forceinline__ float reward(const EnvAction& a, const float (&y)[SD], const float (&yn)[SD]) const {
    if constexpr (KIND == LRX_ENV_CARTPOLE) {
      return 1.0f;
    } else if constexpr (KIND == LRX_ENV_PENDULUM) {
      float u = lrx_clipf(a.f, -c[1], c[1]);
      float cost = yn[0] * yn[0] + 0.1f * (yn[1] * yn[1]) + 0.001f * (u * u);
      return -cost;
    } else if constexpr (KIND == LRX_ENV_MOUNTAINCAR) {
      return -1.0f;
    } else if constexpr (KIND == LRX_ENV_CONTINUOUS_MOUNTAINCAR) {
      const float act = lrx_clipf(a.f, c[0], c[1]);
      return 100.0f * (terminal(y) ? 1.0f : 0.0f) - 0.1f * (act * act);
    } else {
      return terminal(yn) ? 0.0f : -1.0f;
    }
  }

  // ---- terminal: cartpole.py:197-203, pendulum.py:125-126, acrobot.py:272-277
  __device__ __forceinline__ bool terminal(const float (&y)[SD]) const {
    if constexpr (KIND == LRX_ENV_CARTPOLE) {
      const float tt = c[5], xt = c[6];
      bool within = (y[0] >= -xt) && (y[0] <= xt) && (y[2] >= -tt) && (y[2] <= tt);
      return !within;
    } else if constexpr (KIND == LRX_ENV_PENDULUM) {
      return false;
    } else if constexpr (KIND == LRX_ENV_MOUNTAINCAR) {
      return (y[0] >= c[3]) && (y[1] >= c[4]);  // mountain_car.py:182-186
    } else if constexpr (KIND == LRX_ENV_CONTINUOUS_MOUNTAINCAR) {
      return (y[0] >= c[5]) && (y[1] >= c[6]);  // continuous_mountain_car.py:170-174
    } else {
      return (-cosf(y[0]) - cosf(y[0] + y[1])) > 1.0f;
    }
  }

  // ---- TimeLimit.truncate: wrapper/misc.py:193-195 (bare env: base_classic_control.py:74-75)
  __device__ __forceinline__ bool truncate(int step_count) const {
    return max_steps > 0 && step_count >= max_steps;
  }

  // ---- env.initial from uniforms in [0,1): cartpole.py:157-160, pendulum.py:84-87,
  // acrobot.py:162-165.  jax.random.uniform = max(minval, u*(maxval-minval)+minval).
  __device__ __forceinline__ void initial(const float (&u)[SD], float (&y)[SD]) const {
    if constexpr (KIND == LRX_ENV_CARTPOLE) {
#pragma unroll
      for (int i = 0; i < 4; ++i) y[i] = fmaxf(-0.05f, __fadd_rn(__fmul_rn(u[i], 0.05f - (-0.05f)), -0.05f));
    } else if constexpr (KIND == LRX_ENV_PENDULUM) {
      const float PI_F = 3.14159265358979323846f;
      y[0] = fmaxf(-PI_F, __fadd_rn(__fmul_rn(u[0], PI_F - (-PI_F)), -PI_F));
      y[1] = fmaxf(-1.0f, __fadd_rn(__fmul_rn(u[1], 2.0f), -1.0f));
    } else if constexpr (KIND == LRX_ENV_MOUNTAINCAR || KIND == LRX_ENV_CONTINUOUS_MOUNTAINCAR) {
      // mountain_car.py:145-149, continuous_mountain_car.py:132-136: x ~ U(-0.6, -0.4), v = 0
      y[0] = fmaxf(-0.6f, __fadd_rn(__fmul_rn(u[0], -0.4f - (-0.6f)), -0.6f));
      y[1] = 0.0f;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) y[i] = fmaxf(-0.1f, __fadd_rn(__fmul_rn(u[i], 0.1f - (-0.1f)), -0.1f));
    }
  }
};
