/* lerax_b200 -- C ABI of the B200-native backend for Lerax's on-policy PPO hot path.
 *
 * The reference (RunnersNum40/lerax 0.0.6) is pure Python/JAX and has NO plugin, operator
 * or FFI interface for this path: the seam is its Python API.  The entry points below are
 * what a jax.ffi / ctypes binding of that path binds; each cites the reference function it
 * replaces (paths relative to /root/reference/src/lerax).  INTEGRATION.md shows the
 * reference-side stubs.
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer unless the name ends in `_host`; the caller owns all
 *     buffers, the library never allocates or frees on behalf of the caller (workspaces are
 *     caller-allocated, sized by the *_workspace_bytes queries).
 *   - Functions only ENQUEUE work on `stream` (a cudaStream_t passed as void*); none of them
 *     synchronises the host, so they are legal inside an XLA FFI handler or a CUDA graph
 *     capture.  They are re-entrant and keep no global state besides the thread-local
 *     error string and launch counter.  The ONE exception to "never allocates / never synchronises" is the
 *     lrx_p2p_alloc / lrx_p2p_open / close / free family: CUDA IPC can only share memory the exporting process
 *     obtained from cudaMalloc, so those four set-up calls allocate and synchronise (once per process, outside any
 *     capture).  Test hooks (kernel-selection toggles, tensor-core self-tests) live in lerax_b200_debug.h.
 *   - Return value: LRX_OK (0) or a negative LRX_ERR_*; lrx_last_error() gives the message.
 *   - Layouts: the rollout buffer is TIME-MAJOR [T][N] (env index fastest) so that lanes of a
 *     warp touch consecutive envs; the reference's logical layout is [N][T] with flat index
 *     i = n*T + t (buffer/base_buffer.py:38-62).  Minibatch index arrays (`idx`) are in the
 *     REFERENCE order i = n*T + t; the translation to t*N + n happens inside the library.
 *   - All floats are IEEE fp32 (the reference runs JAX with x64 disabled), ints int32,
 *     dones uint8 (0/1).
 *   - There is no CPU fallback and no multi-backend dispatch: a missing device or an
 *     unsupported env/policy combination is an error, never a silent slow path.
 */
#ifndef LERAX_B200_H
#define LERAX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LRX_VERSION 100 /* 0.1.0 */

#define LRX_OK 0
#define LRX_ERR_INVALID_ARG (-1)
#define LRX_ERR_UNSUPPORTED (-2)
#define LRX_ERR_CUDA (-3)

typedef void* lrx_stream_t; /* cudaStream_t */

/* ------------------------------------------------------------------ environments
 * env/classic_control/{cartpole,pendulum,acrobot,mountain_car,continuous_mountain_car}.py.  Physical parameters are RUNTIME
 * values (curriculum callbacks rewrite them between iterations, curriculum/scheduled.py:133-138).
 * p[] per kind:
 *   CARTPOLE : 0 gravity, 1 cart_mass, 2 pole_mass, 3 half_length, 4 force_mag,
 *              5 theta_threshold_radians, 6 x_threshold                 (cartpole.py:110-122)
 *   PENDULUM : 0 max_speed, 1 max_torque, 2 g, 3 m, 4 l                   (pendulum.py:51-62)
 *   ACROBOT  : 0 gravity, 1 link_length_1, 2 link_length_2, 3 link_mass_1, 4 link_mass_2,
 *              5 link_com_pos_1, 6 link_com_pos_2, 7 link_moi, 8 max_vel_1, 9 max_vel_2,
 *              10..12 torques[0..2]                                       (acrobot.py:114-132)
 *   MOUNTAINCAR : 0 min_position, 1 max_position, 2 max_speed, 3 goal_position, 4 goal_velocity,
 *              5 force, 6 gravity                                         (mountain_car.py:106-129)
 *   CONTINUOUS_MOUNTAINCAR : 0 min_action, 1 max_action, 2 min_position, 3 max_position, 4 max_speed,
 *              5 goal_position, 6 goal_velocity, 7 power                  (continuous_mountain_car.py:91-113)
 * max_episode_steps > 0 applies wrapper/misc.py:115-206 (TimeLimit): truncate when the
 * per-episode step counter reaches it; 0 = bare env, never truncates
 * (base_classic_control.py:74-75). */
#define LRX_ENV_CARTPOLE 0
#define LRX_ENV_PENDULUM 1
#define LRX_ENV_ACROBOT 2
#define LRX_ENV_MOUNTAINCAR 3
#define LRX_ENV_CONTINUOUS_MOUNTAINCAR 4
#define LRX_ENV_MAX_PARAMS 16

typedef struct LrxEnvDesc {
  int32_t kind;
  int32_t max_episode_steps;
  float dt;
  float p[LRX_ENV_MAX_PARAMS];
} LrxEnvDesc;

/* Per-env carried state, struct-of-arrays: y[k][N] (k = 4, 2, 4, 2, 2), t[N], step_count[N]
 * (CartPoleState/PendulumState/AcrobotState .y/.t and TimeLimitState.step_count). */
typedef struct LrxEnvState {
  float* y;
  float* t;
  int32_t* step_count;
} LrxEnvState;

/* ------------------------------------------------------------------------ policy
 * policy/actor_critic/mlp.py:64-110 + policy/actor.py:194-261.  Three Linear chains
 * (0 encoder, 1 value_head, 2 action_head.mlp followed by action_dist.mapping) given by
 * their layer widths dims[c][0..n_layers[c]] and a ReLU flag per Linear.  The flat
 * parameter vector is in Equinox leaf order: per chain, per Linear, weight[out][in]
 * row-major then bias[out]; then log_std (Box only).  n_params must equal
 * lrx_policy_num_params(). */
#define LRX_MAX_LAYERS 8
#define LRX_MAX_NOUT 8

typedef struct LrxPolicyDesc {
  int32_t obs_dim;
  int32_t n_out;  /* logits (Discrete) or 1 (scalar Box) */
  int32_t is_box; /* 1: Normal(loc, exp(log_std)); 0: Categorical(logits) */
  int32_t n_params;
  int32_t n_layers[3];
  int32_t dims[3][LRX_MAX_LAYERS + 1];
  int32_t relu[3][LRX_MAX_LAYERS];
} LrxPolicyDesc;

int lrx_policy_num_params(const LrxPolicyDesc* pol);

/* --------------------------------------------------------------------------- RNG
 * Counter-based Philox4x32-10: key = seed, counter = (env_offset + n, step_offset + s,
 * stream, 0); stream 0 = action noise (Gumbel per logit, or one N(0,1) by Box-Muller),
 * stream 1 = reset noise (k uniforms).  Replaces the jr.split key tree of
 * algorithm/ppo.py:209-219 (RNG streams are not part of the parity contract). */
typedef struct LrxRng {
  uint64_t seed;
  uint32_t env_offset;
  uint32_t step_offset;
} LrxRng;

/* Injected noise for parity runs ("same action/noise sequences"): when non-NULL the
 * kernels read noise from here instead of generating it.
 *   action_noise  [T][N][n_out] Gumbel (Discrete) or [T][N] eps (Box)
 *   reset_uniform [T][N][k]     uniforms in [0,1) consumed by env.initial when done */
typedef struct LrxReplay {
  const float* action_noise;
  const float* reset_uniform;
} LrxReplay;

/* Rollout buffer planes, all [T][N] (obs [T][N][obs_dim]); RolloutBuffer fields of
 * buffer/rollout.py:21-30.  act is int32 (Discrete) or float (Box).  The *_trace
 * pointers are optional (NULL to skip): pre-transition state y [T][k][N], its time
 * t [T][N], and the post-transition PRE-reset state [T][k][N] -- used by the
 * teacher-forced parity tests. */
typedef struct LrxRolloutOut {
  float* obs;
  void* act;
  float* rew;
  uint8_t* done;
  float* logp;
  float* val;
  float* last_value; /* [N] : V(obs(state after T steps)), ppo.py:311-312 */
  float* y_trace;
  float* t_trace;
  float* y_next_trace;
  /* Optional scratch for policies with NON-default widths (lrx_rollout_workspace_bytes; NULL = none): with it a rollout
   * of >= 512 envs runs step by step as batched GEMMs over the envs instead of one lane per env. */
  void* workspace;
  size_t workspace_bytes;
} LrxRolloutOut;

/* Bytes of LrxRolloutOut.workspace for N envs (0 on invalid arguments). */
size_t lrx_rollout_workspace_bytes(const LrxPolicyDesc* pol, int32_t N);

/* env.initial for N envs (PPOStepState.initial, algorithm/ppo.py:45-58,331-333).
 * uniform: optional injected draws [N][k] in [0,1); NULL -> Philox stream 1 at
 * step index rng.step_offset. */
int lrx_env_reset(const LrxEnvDesc* env, LrxEnvState state, int32_t N, LrxRng rng,
                  const float* uniform, lrx_stream_t stream);

/* One batched env.transition + reward + terminal + truncate + observation, WITHOUT the
 * reset-on-done (env/classic_control/base_classic_control.py:49-72 and the env files).
 * state is updated in place.  action [N] int32/float; outputs may be NULL. */
int lrx_env_step(const LrxEnvDesc* env, LrxEnvState state, const void* action, float* reward,
                 uint8_t* terminal, uint8_t* truncated, float* obs_next, int32_t N,
                 lrx_stream_t stream);

/* env.observation and env.terminal for N envs: obs [N][obs_dim], terminal [N] (either may
 * be NULL). */
int lrx_env_observe(const LrxEnvDesc* env, LrxEnvState state, float* obs, uint8_t* terminal,
                    int32_t N, lrx_stream_t stream);

/* Batched network forward (policy.__call__/value/evaluate_action building block,
 * policy/actor_critic/mlp.py:115-178): obs [B][obs_dim] -> value [B], head [B][n_out]
 * (raw logits or loc).  Either output may be NULL. */
int lrx_policy_forward(const LrxPolicyDesc* pol, const float* params, const float* obs,
                       float* value, float* head, int32_t B, lrx_stream_t stream);

/* The fused persistent rollout: T steps of PPO.step for N envs + last_value
 * (algorithm/ppo.py:199-316 under the vmap of :360-368).  state is carried in/out. */
int lrx_rollout(const LrxEnvDesc* env, const LrxPolicyDesc* pol, const float* params,
                LrxEnvState state, LrxRng rng, const LrxReplay* replay, LrxRolloutOut out,
                int32_t N, int32_t T, float gamma, lrx_stream_t stream);

/* ---------------------------------------------------------------------------- multi-GPU gradient exchange
 * The reference has no collective (single-device vmap); on G GPUs the data-parallel update sums
 * gradbuf over the ranks between backward and clip + Adam.  These entry points do that
 * sum over NVLink peer memory instead of a collective library call: every rank allocates an exchange
 * buffer, hands its 64-byte IPC handle to the peers (any host channel), opens theirs, and passes a
 * DEVICE array of the G buffer addresses (own buffer at index `rank`) to lrx_p2p_allreduce.
 * Per minibatch: lrx_ppo_grad writes into lrx_p2p_slot(own, n, seq); lrx_p2p_allreduce(seq) leaves the
 * rank-ordered sum in `out` (bit-identical on every rank); lrx_ppo_apply consumes `out`.
 * seq starts at 1 and increases by 1 per exchange on every rank. */
#define LRX_P2P_HANDLE_BYTES 64
size_t lrx_p2p_buffer_bytes(int32_t n_floats);
int lrx_p2p_alloc(int32_t n_floats, void** buffer, unsigned char* ipc_handle /* [LRX_P2P_HANDLE_BYTES] */);
int lrx_p2p_open(const unsigned char* ipc_handle, void** buffer);
int lrx_p2p_close(void* peer_buffer);
int lrx_p2p_free(void* own_buffer);
float* lrx_p2p_slot(void* own_buffer, int32_t n_floats, uint32_t seq);
/* error_flag: device int set to 1 if a peer did not arrive within ~10 s (instead of hanging). */
int lrx_p2p_allreduce(void* const* peer_buffers, int32_t rank, int32_t world, int32_t n_floats, float* out,
                      uint32_t seq, int32_t* error_flag, lrx_stream_t stream);

/* ---------------------------------------------------------------------------- built-in on_step states
 * LoggingCallbackStepState (callback/logging/callback.py:90-176) of every env, advanced over
 * the T steps of a rollout from the reward / done streams the rollout kernel wrote: replaces
 * LoggingCallback.on_step (:477-480) traced into the per-step scan.  All arrays are [N] device
 * arrays; they carry the state in and out (zero-initialised = LoggingCallbackStepState.initial()).
 * ProgressBarCallbackStepState.steps (callback/progress_bar.py:177-181) is `step`. */
typedef struct {
  int64_t* step;            /* env steps seen */
  float* episode_return;    /* running return of the current episode */
  int32_t* episode_length;
  uint8_t* episode_done;    /* the previous step ended an episode */
  float* average_return;    /* EMA over finished episodes, weight alpha on the newest */
  float* average_length;
} LrxEpisodeStats;
/* layout: LRX_LAYOUT_TN / LRX_LAYOUT_NT as for lrx_gae. */
int lrx_episode_stats(const float* rew, const uint8_t* done, const LrxEpisodeStats* st, int32_t N, int32_t T,
                      float alpha, int32_t layout, lrx_stream_t stream);

/* ---------------------------------------------------------------------------- GAE
 * advantage/gae.py:36-66 via buffer/rollout.py:64-79.  layout LRX_LAYOUT_TN: arrays are
 * [T][N] (kernel-native); LRX_LAYOUT_NT: [N][T] (reference-native, T contiguous).
 * ev_sums: optional double[LRX_GAE_EV_DOUBLES], ZERO-INITIALISED ONCE by the caller when it is allocated.
 * ev_sums[0..3] receive sum(ret), sum(ret^2), sum(adv), sum(adv^2) (adv = ret - val) for explained_variance
 * (algorithm/ppo.py:523-528); the rest is per-block scratch and an arrival counter that the kernel re-arms itself,
 * so the sums are added in a fixed order (bitwise reproducible) and no memset precedes the kernel. */
#define LRX_LAYOUT_TN 0
#define LRX_LAYOUT_NT 1
#define LRX_GAE_EV_BLOCKS 2048
#define LRX_GAE_EV_DOUBLES (4 + 4 * LRX_GAE_EV_BLOCKS + 4)
int lrx_gae(const float* rew, const float* val, const uint8_t* done, const float* last_value,
            float* adv, float* ret, int32_t N, int32_t T, float gamma, float lam, int32_t layout,
            double* ev_sums, lrx_stream_t stream);
/* The reference's two other estimators, same arrays and layouts: n_step == 0 is BootstrappedReturn
 * (advantage/bootstrapped.py:12-79: G_t = r_t + gamma (1 - done_t) G_{t+1}, G_T = last_value), n_step >= 1 is
 * NStepReturn(n) (advantage/n_step.py:30-70).  adv = ret - val. */
int lrx_returns(const float* rew, const float* val, const uint8_t* done, const float* last_value, float* adv,
                float* ret, int32_t N, int32_t T, float gamma, int32_t n_step, int32_t layout, lrx_stream_t stream);

/* ------------------------------------------------------------------- permutation
 * buffer/base_buffer.py:85-88: jr.permutation(key, N*T).  out[n] receives a pseudo-random
 * permutation of [0, n) keyed by `seed` (keyed Feistel bijection with cycle walking, one pass;
 * the RNG stream is not part of the parity contract). */
int lrx_permutation(int32_t* out, int64_t n, uint64_t seed, lrx_stream_t stream);

/* --------------------------------------------------------------------- PPO update
 * algorithm/ppo.py:396-561 + optax chain(clip_by_global_norm, adam) (ppo.py:192-194). */
typedef struct LrxPpoHyper {
  float clip_coefficient;
  float value_loss_coefficient;
  float entropy_loss_coefficient;
  float max_grad_norm;
  float learning_rate;
  float adam_b1, adam_b2, adam_eps;
  int32_t clip_value_loss;
  int32_t normalize_advantages;
  int32_t surrogate; /* 0: PPO clipped ratio (algorithm/ppo.py:429-434); 1: -mean(log_prob * advantage), the
                        policy loss of A2C and REINFORCE (algorithm/a2c.py:401, algorithm/reinforce.py:391) */
} LrxPpoHyper;

/* Flat view of the rollout buffer the update gathers from; planes are [T][N]. */
typedef struct LrxBatchView {
  const float* obs;  /* [T][N][obs_dim] */
  const void* act;   /* [T][N] */
  const float* logp; /* [T][N] */
  const float* val;  /* [T][N] */
  const float* ret;  /* [T][N] */
  const float* adv;  /* [T][N] */
  int32_t N, T;
  /* Optional (NULL = gather from the planes above): row-major copy of the buffer written by lrx_ppo_pack_rows, one
   * LRX_PACKED_FLOATS-float record per row in REFERENCE order i = n*T + t.  A permuted minibatch row then costs one
   * 64-byte line instead of one 32-byte sector per plane (13.9x the algorithmic bytes, profiles/r1_ppo_grad_fused.md). */
  const float* packed;
} LrxBatchView;

/* Record layout of LrxBatchView.packed: floats 0..7 observation (zero padded), 8 action bits (int32 / float), 9 old
 * log-prob, 10 old value, 11 return, 12 advantage, 13..15 zero.  Only policies with obs_dim <= 8 use it. */
#define LRX_PACKED_FLOATS 16
/* [N*T][LRX_PACKED_FLOATS] from the planes of `buf` (buf->packed is ignored); one pass, 41 B read + 64 B written per
 * row.  The flatten of buffer/base_buffer.py:38-62 made physical: record i = n*T + t. */
int lrx_ppo_pack_rows(const LrxBatchView* buf, int32_t obs_dim, int32_t is_box, float* packed, lrx_stream_t stream);

#define LRX_PPO_NSTATS 8 /* approx_kl, loss, policy_loss, value_loss, entropy_loss, grad_norm, nonfinite, 0 */

/* Bytes of caller-allocated workspace for minibatches of up to B rows. */
size_t lrx_ppo_workspace_bytes(const LrxPolicyDesc* pol, int64_t B);

/* Per-minibatch advantage sums for the normalisation of ppo.py:424-427:
 * sums[b] = (sum A, sum A^2) over idx[b][0..B) in double.  One launch for a whole epoch
 * so that a multi-GPU caller can all-reduce all of them in ONE collective. */
int lrx_ppo_adv_sums(const float* adv, int32_t N, int32_t T, const int32_t* idx,
                     int32_t num_batches, int64_t B, double* sums, lrx_stream_t stream);

/* Loss forward + backward for ONE minibatch share of B rows (ppo_loss + grad,
 * ppo.py:396-469).  B_global is the size of the whole minibatch across ranks and adv_sums
 * its (all-reduced) advantage sums.  gradbuf[n_params + LRX_PPO_NSTATS] receives this
 * share's contribution to the gradient of the MEAN loss and to the stat sums, so that a
 * plain SUM all-reduce over ranks yields the global quantities. */
int lrx_ppo_grad(const LrxPolicyDesc* pol, const float* params, const LrxBatchView* buf,
                 const int32_t* idx, int64_t B, int64_t B_global, const double* adv_sums,
                 const LrxPpoHyper* h, float* gradbuf, void* workspace, size_t workspace_bytes,
                 lrx_stream_t stream);

/* clip_by_global_norm + Adam on the (all-reduced) gradbuf (train_batch, ppo.py:471-492);
 * writes this minibatch's LRX_PPO_NSTATS stats and ORs the non-finite flag of
 * ppo.py:413-417 into *nonfinite_flag. */
int lrx_ppo_apply(const LrxPolicyDesc* pol, float* params, float* adam_m, float* adam_v,
                  int32_t* adam_count, const float* gradbuf, const LrxPpoHyper* h,
                  float* stats_out, int32_t* nonfinite_flag, lrx_stream_t stream);

/* lrx_p2p_allreduce + lrx_ppo_apply in ONE cooperative launch: every block sums its 256 gradient elements over the
 * ranks' exchange slots, the global norm is assembled at a grid barrier, then clip + Adam (train_batch,
 * ppo.py:471-492).  The exchange buffers hold n_params + LRX_PPO_NSTATS floats per slot. */
size_t lrx_ppo_apply_p2p_scratch_bytes(const LrxPolicyDesc* pol);
int lrx_ppo_apply_p2p(const LrxPolicyDesc* pol, float* params, float* adam_m, float* adam_v, int32_t* adam_count,
                      void* const* peer_buffers, int32_t rank, int32_t world, uint32_t seq, const LrxPpoHyper* h,
                      float* stats_out, int32_t* nonfinite_flag, int32_t* error_flag, void* scratch,
                      size_t scratch_bytes, lrx_stream_t stream);

/* lrx_ppo_grad into slot `seq & 1` of this rank's exchange buffer (lrx_p2p_alloc), with the slot published (flag = seq) by
 * the last block of the partial reduction: the flag reaches the peers while lrx_ppo_apply_p2p is being launched. */
int lrx_ppo_grad_p2p(const LrxPolicyDesc* pol, const float* params, const LrxBatchView* buf, const int32_t* idx, int64_t B,
                     int64_t B_global, const double* adv_sums, const LrxPpoHyper* h, void* own_buffer, uint32_t seq,
                     void* workspace, size_t workspace_bytes, lrx_stream_t stream);

/* One epoch of the data-parallel update in ONE call: for each of the num_batches rows of idx, lrx_ppo_grad_p2p and
 * lrx_ppo_apply_p2p with sequence numbers derived from seq64 + 1 .. seq64 + num_batches (train_epoch, ppo.py:494-521, with the
 * gradient SUM over the ranks inside).  adv_sums [num_batches][2]: the all-reduced advantage sums of the epoch;
 * learning_rates_host: optional per-minibatch rates (a schedule), else h->learning_rate. */
int lrx_ppo_update_p2p(const LrxPolicyDesc* pol, float* params, float* adam_m, float* adam_v, int32_t* adam_count,
                       const LrxBatchView* buf, const int32_t* idx, int32_t num_batches, int64_t B, int64_t B_global,
                       const double* adv_sums, const LrxPpoHyper* h, const float* learning_rates_host,
                       void* const* peer_buffers, void* own_buffer, int32_t rank, int32_t world, uint64_t seq64,
                       float* stats_out, int32_t* nonfinite_flag, int32_t* error_flag, void* scratch, size_t scratch_bytes,
                       void* workspace, size_t workspace_bytes, lrx_stream_t stream);

/* One minibatch of the multi-GPU update in two launches (default-width policies): lrx_ppo_grad's gradient kernel, then ONE
 * cooperative kernel that reduces the per-CTA partials into this rank's exchange slot `seq & 1`, publishes ONE flag per rank
 * (the last block to finish the reduction), waits for the peers (block 0 polls their flags over NVLink and relays the
 * arrival through a flag in local memory), sums the slots in rank order and applies clip + Adam (backward -> gradient SUM
 * over the ranks -> train_batch, ppo.py:469-492).  `peer_buffers` is the DEVICE array of the G exchange buffers
 * (lrx_p2p_alloc / lrx_p2p_open), `own_buffer` this rank's.  *handled = 0 and nothing is enqueued for policies the fused
 * gradient kernels do not cover (run lrx_ppo_grad_p2p + lrx_ppo_apply_p2p instead).  Environment: LRX_P2P_FUSED=1 one flag
 * per block, =3 push variant (both measured slower, profiles/r2_configs.md). */
int lrx_ppo_grad_apply_p2p(const LrxPolicyDesc* pol, float* params, float* adam_m, float* adam_v, int32_t* adam_count,
                           const LrxBatchView* buf, const int32_t* idx, int64_t B, int64_t B_global,
                           const double* adv_sums, const LrxPpoHyper* h, void* const* peer_buffers, void* own_buffer,
                           int32_t rank, int32_t world, uint32_t seq, float* stats_out, int32_t* nonfinite_flag,
                           int32_t* error_flag, void* workspace, size_t workspace_bytes, lrx_stream_t stream,
                           int32_t* handled);

/* One epoch on one GPU: for each of num_batches rows of idx: adv sums, grad, apply
 * (train_epoch, ppo.py:494-521).  stats_out [num_batches][LRX_PPO_NSTATS]. */
int lrx_ppo_update(const LrxPolicyDesc* pol, float* params, float* adam_m, float* adam_v,
                   int32_t* adam_count, const LrxBatchView* buf, const int32_t* idx,
                   int32_t num_batches, int64_t B, const LrxPpoHyper* h, float* stats_out,
                   int32_t* nonfinite_flag, void* workspace, size_t workspace_bytes,
                   lrx_stream_t stream);

/* The same with a learning-rate SCHEDULE (ppo.py:176,192-194: optax.inject_hyperparams evaluates a callable
 * learning_rate at its own update count): learning_rates_host is a HOST array of num_batches rates, one per minibatch
 * (NULL: h->learning_rate for all of them). */
int lrx_ppo_update_schedule(const LrxPolicyDesc* pol, float* params, float* adam_m, float* adam_v,
                            int32_t* adam_count, const LrxBatchView* buf, const int32_t* idx,
                            int32_t num_batches, int64_t B, const LrxPpoHyper* h, const float* learning_rates_host,
                            float* stats_out, int32_t* nonfinite_flag, void* workspace, size_t workspace_bytes,
                            lrx_stream_t stream);

/* ------------------------------------------------------------------------- misc */
const char* lrx_last_error(void);
int lrx_version(void);
/* Number of kernel launches issued by this library on the calling thread since the last
 * call with reset != 0 (bench.py's "gpu_launches"). */
int64_t lrx_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* LERAX_B200_H */
