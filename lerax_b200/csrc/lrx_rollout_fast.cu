// Fused persistent rollout for the default MLPActorCriticPolicy widths on the tensor cores.
//
// Reference: algorithm/ppo.py:199-316 (PPO.step / collect_rollout) under the vmap of :360-368;
// policy/actor_critic/mlp.py:133-157; the env functors of lrx_env.cuh.  Same contract as the
// generic kernel in lrx_rollout.cu (which keeps serving non-default widths); tried first.
//
// One CTA owns 128 envs for all T steps (one UMMA M = 128 tile: env i of the CTA = TMEM lane i).
// The Linear weights sit in shared memory as bf16 pieces (lrx_tc.cuh, fp32 emulation with six
// bf16 products per fp32 product); every policy layer is one group of tcgen05.mma whose fp32
// result is pulled from TMEM by the env's thread (warp w: TMEM quadrant w & 3, column group
// w >> 2), biased / ReLU'd and written back to shared memory as the next layer's operand.  The
// env state, the Tsit5 step, sampling (Philox4x32-10), reward / done / reset stay in registers
// of the 128 threads of column group 0, which also stream the buffer row to HBM (coalesced over
// envs).  Per env-step: 5 MMA round trips (enc0, enc1, enc2, both heads' layer 1, both heads'
// layer 2); the value bootstrap of a truncated step (ppo.py:248-259) re-runs the value path for
// the whole tile only on steps where some env of the CTA truncates.
#include "lrx_env.cuh"
#include "lrx_policy.cuh"
#include "lrx_tc.cuh"

namespace {

constexpr int ENVS = 128;      // envs per CTA = UMMA M
constexpr int THREADS = 512;   // 16 warps
constexpr int ISSUERS = 2;     // warps 0 and 1 issue (value / action head GEMMs in parallel)
constexpr uint32_t LRX_STREAM_INIT_ = 2;

constexpr uint32_t PW16 = 64 * 16 * 2, PW64 = 64 * 64 * 2;          // weight pieces ([64][16], [64][64] / [16][64] = PW16)
constexpr uint32_t PX16 = ENVS * 16 * 2, PX64 = ENVS * 64 * 2;      // activation pieces
// Composite weights across the policy's two linear-linear seams (the same identity as lrx_update_tc128c.cu): the encoder's
// 16-wide output has no activation, so both heads' first layers read h2 through Wcv = Wv0 W2 / Wca = Wa0 W2 (64 x 64), and
// action_dist.mapping follows the action head's linear output layer, so the logits are (Wmap Wa1) xa + (Wmap ba1 + bmap).
// Four MMA round trips per env-step instead of five, no 16 -> n_out mapping on the FMA pipe.  The composites are formed
// once per launch in fp32 (one rounding of a 16-term dot product per weight, ~1e-7 relative: the rollout's contract is
// 1e-5) and then split exactly into three pieces like every other operand.
constexpr uint32_t OFF_W0 = 0;
constexpr uint32_t OFF_W1 = OFF_W0 + 3 * PW16;
constexpr uint32_t OFF_WCV = OFF_W1 + 3 * PW64;
constexpr uint32_t OFF_WCA = OFF_WCV + 3 * PW64;
constexpr uint32_t OFF_WV1 = OFF_WCA + 3 * PW64;
constexpr uint32_t OFF_WCM = OFF_WV1 + 3 * PW64;    // [16][64]: rows >= n_out zero
constexpr uint32_t OFF_XOBS = OFF_WCM + 3 * PW16;
constexpr uint32_t OFF_XH1 = OFF_XOBS + 3 * PX16;   // encoder layer 1, then value head layer 0
constexpr uint32_t OFF_XH2 = OFF_XH1 + 3 * PX64;    // encoder layer 2 (h2), then action head layer 0
constexpr uint32_t OFF_F32 = OFF_XH2 + 3 * PX64;
constexpr uint32_t OFF_VPART = OFF_F32 + tc::F_RESERVE * 4;   // [4][128] partial value dot products
constexpr uint32_t OFF_MISC = OFF_VPART + 4 * ENVS * 4;
constexpr uint32_t SMEM_BYTES = OFF_MISC + 64;
static_assert(SMEM_BYTES <= 227 * 1024, "fused rollout kernel exceeds shared memory");

constexpr uint32_t C_D0 = 0, C_D1 = 64, TMEM_COLS = 128;

struct RolloutArgs {
  LrxEnvDesc env;
  const float* params;
  LrxEnvState st;
  LrxRng rng;
  LrxReplay rp;
  LrxRolloutOut out;
  int N, T;
  float gamma;
  int n_out;
  tc::LayerOffsets lo;
};

template <int KIND>
__global__ void __launch_bounds__(THREADS, 1) rollout_tc_kernel(const __grid_constant__ RolloutArgs a) {
  using E = Env<KIND>;
  constexpr int SD = E::SD, OD = E::OD;
  constexpr bool BOX = EnvDims<KIND>::BOX;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = umma::warp_idx_sync(), lane = tid & 31;
  const int q = warp & 3, g = warp >> 2;
  const bool g0 = (g == 0);
  const int m = q * 32 + lane;  // env of the CTA owned by this thread = TMEM lane
  float* fp = reinterpret_cast<float*>(smem + OFF_F32);
  float* vpart = reinterpret_cast<float*>(smem + OFF_VPART);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_MISC);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + OFF_MISC + 8);

  const tc::Mat W0{smem + OFF_W0, 64, PW16}, W1{smem + OFF_W1, 64, PW64}, WV1{smem + OFF_WV1, 64, PW64};
  const tc::Mat WCV{smem + OFF_WCV, 64, PW64}, WCA{smem + OFF_WCA, 64, PW64}, WCM{smem + OFF_WCM, 16, PW16};
  const tc::Mat XOBS{smem + OFF_XOBS, ENVS, PX16}, XH1{smem + OFF_XH1, ENVS, PX64}, XH2{smem + OFF_XH2, ENVS, PX64};

  tc::stage_weight(a.params + a.lo.w[tc::L_ENC0], 64, OD, 16, W0, tid, THREADS);
  tc::stage_weight(a.params + a.lo.w[tc::L_ENC1], 64, 64, 64, W1, tid, THREADS);
  tc::stage_weight(a.params + a.lo.w[tc::L_V1], 64, 64, 64, WV1, tid, THREADS);
  tc::stage_f32(a.params, a.lo, a.n_out, OD, fp, tid, THREADS);
  {
    // fp32 sources of the composites in the (still unused) XH1 region:
    //   W2 [16][64] | Wv0 [64][16] | Wa0 [64][16] | Wa1 [16][64] | Wmap [8][16] (rows >= n_out zero)
    float* cs = reinterpret_cast<float*>(smem + OFF_XH1);
    constexpr int CS_W2 = 0, CS_WV0 = 1024, CS_WA0 = 2048, CS_WA1 = 3072, CS_WMAP = 4096, CS_N = 4224;
    for (int i = tid; i < CS_N; i += THREADS) {
      float x;
      if (i < CS_WV0) x = __ldg(a.params + a.lo.w[tc::L_ENC2] + i);
      else if (i < CS_WA0) x = __ldg(a.params + a.lo.w[tc::L_V0] + (i - CS_WV0));
      else if (i < CS_WA1) x = __ldg(a.params + a.lo.w[tc::L_A0] + (i - CS_WA0));
      else if (i < CS_WMAP) x = __ldg(a.params + a.lo.w[tc::L_A1] + (i - CS_WA1));
      else x = (i - CS_WMAP) < a.n_out * 16 ? __ldg(a.params + a.lo.w[tc::L_MAP] + (i - CS_WMAP)) : 0.f;
      cs[i] = x;
    }
    __syncthreads();  // cs, and the biases stage_f32 wrote
    // chunk e = 8 consecutive inputs i of one output o:  Wc[o][i] = sum_k L[o][k] R[k][i]  (k = 0..15 in order)
    for (int e = tid; e < 512 + 512 + 128; e += THREADS) {
      const int mat = e < 512 ? 0 : (e < 1024 ? 1 : 2);
      const int el = e - (mat == 2 ? 1024 : mat * 512);
      const int o = mat == 2 ? (el & 15) : (el & 63), ch = mat == 2 ? (el >> 4) : (el >> 6);
      const float* Lrow = cs + (mat == 0 ? CS_WV0 : (mat == 1 ? CS_WA0 : CS_WMAP)) + o * 16;
      const float* R = cs + (mat == 2 ? CS_WA1 : CS_W2) + ch * 8;
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (mat != 2 || o < LRX_MAX_NOUT) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float l = Lrow[k];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(l, R[k * 64 + j], acc[j]);
        }
      }
      tc::store8(mat == 0 ? WCV : (mat == 1 ? WCA : WCM), ch, o, acc);
    }
    // composite biases, computed into registers first (they replace entries other threads still read)
    float cb = 0.f;
    int cb_slot = -1;
    if (tid < 64 + 64 + LRX_MAX_NOUT) {
      const int mat = tid < 64 ? 0 : (tid < 128 ? 1 : 2), o = tid - (mat == 2 ? 128 : mat * 64);
      const float* Lrow = cs + (mat == 0 ? CS_WV0 : (mat == 1 ? CS_WA0 : CS_WMAP)) + o * 16;
      const float* bb = fp + (mat == 2 ? tc::F_BA1 : tc::F_B2);
#pragma unroll
      for (int k = 0; k < 16; ++k) cb = fmaf(Lrow[k], bb[k], cb);
      cb += fp[(mat == 0 ? tc::F_BV0 : (mat == 1 ? tc::F_BA0 : tc::F_BMAP)) + o];  // F_BMAP entries >= n_out are zero
      if (mat == 2 && o >= a.n_out) cb = 0.f;
      cb_slot = (mat == 0 ? tc::F_BV0 : (mat == 1 ? tc::F_BA0 : tc::F_BMAP)) + o;
    }
    __syncthreads();  // every read of cs and of the plain biases
    if (cb_slot >= 0) fp[cb_slot] = cb;  // F_BV0 / F_BA0 / F_BMAP now hold the composite biases
  }
  if (tid == 0) {
    umma::mbar_init(bar, ISSUERS);
    umma::fence_mbar_init();
  }
  if (warp == 0) umma::tmem_alloc(tslot, TMEM_COLS);
  umma::fence_async_smem();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tb = *tslot;
  const uint32_t tw = umma::tmem_addr(tb, q * 32, 0);
  uint32_t phase = 0;

  auto sync_issue = [&](auto&& issue) {
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    if (warp < ISSUERS) {
      umma::tc_fence_after();
      const bool leader = umma::elect_one();
      issue(leader, warp);
      if (leader) umma::mma_commit(bar);
      __syncwarp();
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1u;
    umma::tc_fence_after();
  };
  // columns [c0, c0+16) of a layer output -> + bias -> optional ReLU -> operand; fn sees the values
  auto epilogue = [&](uint32_t col, int c0, const float* bias, bool relu, const tc::Mat& dst, auto&& fn) {
    float v[16];
    umma::tmem_ld16(tw + col + c0, v);
    umma::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      v[j] += bias[c0 + j];
      if (relu) v[j] = fmaxf(v[j], 0.f);
    }
    fn(v);
    tc::store8(dst, (c0 >> 3), m, v);
    tc::store8(dst, (c0 >> 3) + 1, m, v + 8);
  };
  auto nop = [](float*) {};
  const int cg = g * 16;

  // MLPActorCriticPolicy forward for the tile (mlp.py:133-157): value for every thread of the env's
  // row, raw action-head outputs (logits / loc) in `head` for column group 0 when want_head.
  float head[LRX_MAX_NOUT];
  auto forward = [&](const float (&o)[OD], bool want_head) -> float {
    if (g0) {
      float ob[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) ob[k] = (k < OD) ? o[k < OD ? k : 0] : 0.f;
      tc::store8(XOBS, 0, m, ob);
      tc::store8(XOBS, 1, m, ob + 8);
    }
    sync_issue([&](bool ld, int p) {
      if (p == 0) tc::gemm<false, false>(ld, tb + C_D0, XOBS, W0, 128, 64, 1, 0u);
    });
    epilogue(C_D0, cg, fp + tc::F_B0, true, XH1, nop);
    sync_issue([&](bool ld, int p) {
      if (p == 0) tc::gemm<false, false>(ld, tb + C_D0, XH1, W1, 128, 64, 4, 0u);
    });
    epilogue(C_D0, cg, fp + tc::F_B1, true, XH2, nop);
    // both heads' layer 0 straight from h2 (composites)
    sync_issue([&](bool ld, int p) {
      if (p == 0) tc::gemm<false, false>(ld, tb + C_D0, XH2, WCV, 128, 64, 4, 0u);
      if (p == 1 && want_head) tc::gemm<false, false>(ld, tb + C_D1, XH2, WCA, 128, 64, 4, 0u);
    });
    epilogue(C_D0, cg, fp + tc::F_BV0, true, XH1, nop);
    if (want_head) epilogue(C_D1, cg, fp + tc::F_BA0, true, XH2, nop);  // h2 is dead: both GEMMs above are complete
    sync_issue([&](bool ld, int p) {
      if (p == 0) tc::gemm<false, false>(ld, tb + C_D0, XH1, WV1, 128, 64, 4, 0u);
      if (p == 1 && want_head) tc::gemm<false, false>(ld, tb + C_D1, XH2, WCM, 128, 16, 4, 0u);  // logits / loc
    });
    {
      float v[16];
      umma::tmem_ld16(tw + C_D0 + cg, v);
      umma::tmem_ld_wait();
      float part = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) part = fmaf(fp[tc::F_WV2 + cg + j], fmaxf(v[j] + fp[tc::F_BV1 + cg + j], 0.f), part);
      vpart[g * ENVS + m] = part;
    }
    if (want_head && g0) {
      float lg[8];
      umma::tmem_ld8(tw + C_D1, lg);
      umma::tmem_ld_wait();
#pragma unroll
      for (int o2 = 0; o2 < LRX_MAX_NOUT; ++o2) head[o2] = lg[o2] + fp[tc::F_BMAP + o2];
    }
    umma::tc_fence_before();
    __syncthreads();
    return (((fp[tc::F_BV2] + vpart[m]) + vpart[ENVS + m]) + vpart[2 * ENVS + m]) + vpart[3 * ENVS + m];
  };

  // ------------------------------------------------------------------ env state (column group 0)
  const int N = a.N, T = a.T;
  const int n = blockIdx.x * ENVS + m;
  const bool live = g0 && n < N;
  const int nl = n < N ? n : N - 1;
  E env;
  env.load(a.env);
  float y[SD], t = 0.f;
  int sc = 0;
#pragma unroll
  for (int i = 0; i < SD; ++i) y[i] = 0.f;
  if (g0) {
#pragma unroll
    for (int i = 0; i < SD; ++i) y[i] = a.st.y[(size_t)i * N + nl];
    t = a.st.t[nl];
    sc = a.st.step_count[nl];
  }
  float scale = 1.f, log_scale = 0.f;
  if (BOX) {
    const float log_std = __ldg(a.params + a.lo.log_std);
    scale = expf(log_std);     // actor.py:68-72: Normal(loc, exp(log_std))
    log_scale = logf(scale);   // distreqx takes log(scale)
  }
  const LrxRolloutOut& out = a.out;
  const LrxReplay& rp = a.rp;
  const LrxRng& rng = a.rng;
  constexpr int n_out = EnvDims<KIND>::NA;  // lrx_rollout validated pol->n_out against the env

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
    const float value = forward(o, true);

    EnvAction act;
    act.i = 0; act.f = 0.f;
    float logp = 0.f, r = 0.f;
    float y1[SD], t1 = 0.f;
    int sc1 = 0;
    bool done = false, trunc = false;
    if (g0) {
      if (BOX) {
        const float loc = head[0];
        float eps;
        if (rp.action_noise) {
          eps = rp.action_noise[row];
        } else {
          Philox4 w = lrx_noise_words(rng, n, s, LRX_STREAM_ACTION);
          float u1 = lrx_u32_to_unit(w.x), u2 = lrx_u32_to_unit(w.y);
          eps = sqrtf(-2.0f * logf(u1)) * cosf(6.28318530717958647692f * u2);
        }
        const float a_raw = loc + scale * eps;
        logp = normal_log_prob(loc, scale, log_scale, a_raw);  // log-prob of the UNCLIPPED sample
        act.f = lrx_clipf(a_raw, env.action_low(), env.action_high());  // ppo.py:228-235 (Box bounds of the env)
      } else {
        float z[LRX_MAX_NOUT];
#pragma unroll
        for (int j = 0; j < LRX_MAX_NOUT; ++j) z[j] = (j < n_out) ? head[j] : 0.f;
        log_softmax_n<LRX_MAX_NOUT>(z, n_out);
        Philox4 w{0u, 0u, 0u, 0u};
        if (!rp.action_noise) w = lrx_noise_words(rng, n, s, LRX_STREAM_ACTION);
        uint32_t ww[4] = {w.x, w.y, w.z, w.w};
        float best = -INFINITY;
        int bi = 0;
#pragma unroll
        for (int j = 0; j < LRX_MAX_NOUT; ++j) {
          if (j < n_out) {
            float gm;
            if (rp.action_noise) {
              gm = rp.action_noise[row * n_out + j];
            } else {
              if (j >= 4 && (j & 3) == 0) {
                Philox4 w2 = philox4x32_10(rng.env_offset + n, rng.step_offset + s, LRX_STREAM_ACTION,
                                           (uint32_t)(j >> 2), (uint32_t)rng.seed, (uint32_t)(rng.seed >> 32));
                ww[0] = w2.x; ww[1] = w2.y; ww[2] = w2.z; ww[3] = w2.w;
              }
              gm = -logf(-logf(lrx_u32_to_unit(ww[j & 3])));
            }
            const float v = z[j] + gm;
            if (v > best) { best = v; bi = j; }  // first maximum wins (jnp.argmax)
          }
        }
        act.i = bi;
#pragma unroll
        for (int j = 0; j < LRX_MAX_NOUT; ++j) if (j == bi) logp = z[j];
      }

      // ---- env.transition / reward / terminal / truncate (ppo.py:237-246)
      env.transition(y, t, act, y1, t1);
      sc1 = sc + 1;
      r = env.reward(act, y, y1);
      const bool term = env.terminal(y1);
      trunc = env.truncate(sc1);
      done = term || trunc;
    }
    // ppo.py:248-259: r += gamma * V(obs(next_state)) with the PRE-reset next state; the value path
    // runs for the whole tile, only on steps where some env of the CTA truncates
    if (__syncthreads_or(live && trunc)) {
      float o1[OD];
#pragma unroll
      for (int k = 0; k < OD; ++k) o1[k] = 0.f;
      if (g0) env.observe(y1, o1);
      const float vn = forward(o1, false);
      if (trunc) r = r + a.gamma * vn;
    }
    if (g0) {
      if (out.y_next_trace && live) {
#pragma unroll
        for (int i = 0; i < SD; ++i) out.y_next_trace[((size_t)s * SD + i) * N + n] = y1[i];
      }
      // ---- emit the buffer row (ppo.py:276-288)
      if (live) {
#pragma unroll
        for (int k = 0; k < OD; ++k) out.obs[row * OD + k] = o[k];
        if (BOX) ((float*)out.act)[row] = act.f; else ((int32_t*)out.act)[row] = act.i;
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
  }

  // ---- last_value = V(obs(final state)) (ppo.py:311-312)
  {
    float o[OD];
    env.observe(y, o);
    const float lv = forward(o, false);
    if (live) {
      out.last_value[n] = lv;
#pragma unroll
      for (int i = 0; i < SD; ++i) a.st.y[(size_t)i * N + n] = y[i];
      a.st.t[n] = t;
      a.st.step_count[n] = sc;
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tb, TMEM_COLS);
}

template <int KIND>
int launch_tc(const RolloutArgs& a, cudaStream_t stream) {
  static bool attr_set[64];  // per kernel instantiation and device
  LRX_CUDA_CHECK(lrx_set_max_smem(rollout_tc_kernel<KIND>, (int)SMEM_BYTES, attr_set));
  rollout_tc_kernel<KIND><<<(a.N + ENVS - 1) / ENVS, THREADS, SMEM_BYTES, stream>>>(a);
  return LRX_OK;
}


}  // namespace

int lrx_rollout_fast(const LrxEnvDesc* env, const LrxPolicyDesc* pol, const float* params,
                     LrxEnvState state, LrxRng rng, const LrxReplay* replay, LrxRolloutOut out,
                     int32_t N, int32_t T, float gamma, cudaStream_t stream, bool* handled) {
  *handled = false;
  if (!tc::is_default_policy(*pol) || !lrx_fused_enabled()) return LRX_OK;
  RolloutArgs a{};
  a.env = *env;
  a.params = params;
  a.st = state;
  a.rng = rng;
  a.rp = *replay;
  a.out = out;
  a.N = N;
  a.T = T;
  a.gamma = gamma;
  a.n_out = pol->n_out;
  a.lo = tc::layer_offsets(*pol);
  int rc;
  switch (env->kind) {
    case LRX_ENV_CARTPOLE: rc = launch_tc<LRX_ENV_CARTPOLE>(a, stream); break;
    case LRX_ENV_PENDULUM: rc = launch_tc<LRX_ENV_PENDULUM>(a, stream); break;
    case LRX_ENV_ACROBOT: rc = launch_tc<LRX_ENV_ACROBOT>(a, stream); break;
    case LRX_ENV_MOUNTAINCAR: rc = launch_tc<LRX_ENV_MOUNTAINCAR>(a, stream); break;
    case LRX_ENV_CONTINUOUS_MOUNTAINCAR: rc = launch_tc<LRX_ENV_CONTINUOUS_MOUNTAINCAR>(a, stream); break;
    default: return LRX_OK;
  }
  if (rc) return rc;
  LRX_LAUNCH_CHECK();
  *handled = true;
  return LRX_OK;
}
