// jax.ffi registration shim: the three hot-path entry points of include/lerax_b200.h as XLA FFI handlers, so that the
// reference (RunnersNum40/lerax, pure JAX) can call them from inside its single `eqx.filter_jit` of `learn`
// (algorithm/base_algorithm.py:208-276) with `jax.ffi.ffi_call`.  INTEGRATION.md shows the Python side.
//
// NOT COMPILED IN THIS IMAGE: jaxlib (and its xla/ffi/api/ffi.h) is not installed and cannot be (no wheels, no network).
// Build where JAX is available:
//     g++ -O2 -fPIC -shared -std=c++17 -DLRX_WITH_XLA_FFI -I"$(python -c 'import jax.ffi; print(jax.ffi.include_dir())')" \
//         -I../../include xla_ffi_shim.cc -L../_C -llerax_b200 -Wl,-rpath,'$ORIGIN' -o ../_C/liblerax_b200_xla.so
// Every handler only ENQUEUES on the stream XLA hands it (the C ABI never synchronises or allocates), which is what XLA
// requires of a custom call.  Scratch memory comes from XLA's own allocator (ffi::ScratchAllocator).
//
// Layout contract: the library's buffers are time-major [T][N] (env index fastest).  The reference's arrays are [N][T]
// (vmap axis outermost, algorithm/ppo.py:360-368); the Python wrapper transposes at the boundary (or, better, keeps the
// rollout buffer time-major end to end -- only `flatten_axes`' index order i = n*T + t is observable, and the library
// honours it).
#ifdef LRX_WITH_XLA_FFI

#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <string>

#include "lerax_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error status(int rc, const char* what) {
  if (rc == LRX_OK) return ffi::Error::Success();
  return ffi::Error(rc == LRX_ERR_CUDA ? ffi::ErrorCode::kInternal : ffi::ErrorCode::kInvalidArgument,
                    std::string(what) + ": " + lrx_last_error());
}

// LrxPolicyDesc from the attributes the Python wrapper derives from MLPActorCriticPolicy's constructor arguments
// (policy/actor_critic/mlp.py:64-110): dims = the three chains' layer widths, concatenated; n_layers = Linears per chain.
LrxPolicyDesc make_policy(int32_t obs_dim, int32_t n_out, bool is_box, ffi::Span<const int32_t> n_layers,
                          ffi::Span<const int32_t> dims) {
  LrxPolicyDesc p{};
  p.obs_dim = obs_dim;
  p.n_out = n_out;
  p.is_box = is_box ? 1 : 0;
  size_t k = 0;
  for (int c = 0; c < 3; ++c) {
    p.n_layers[c] = n_layers[c];
    for (int i = 0; i <= n_layers[c]; ++i) p.dims[c][i] = dims[k++];
    // eqx.nn.MLP: activation after every Linear but the last; the action head's MLP output feeds `mapping` linearly
    for (int i = 0; i < n_layers[c]; ++i) p.relu[c][i] = (c == 2) ? (i < n_layers[c] - 2) : (i < n_layers[c] - 1);
  }
  p.n_params = lrx_policy_num_params(&p);
  return p;
}

LrxEnvDesc make_env(int32_t kind, int32_t max_episode_steps, float dt, ffi::Span<const float> phys) {
  LrxEnvDesc e{};
  e.kind = kind;
  e.max_episode_steps = max_episode_steps;
  e.dt = dt;
  for (size_t i = 0; i < phys.size() && i < LRX_ENV_MAX_PARAMS; ++i) e.p[i] = phys[i];
  return e;
}

// ------------------------------------------------------------------------------------------------- lrx_gae
// advantage/gae.py:36-66 via buffer/rollout.py:64-79.
ffi::Error GaeImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> rew, ffi::Buffer<ffi::F32> val, ffi::Buffer<ffi::U8> done,
                   ffi::Buffer<ffi::F32> last_value, float gamma, float lam, int32_t layout,
                   ffi::ResultBuffer<ffi::F32> adv, ffi::ResultBuffer<ffi::F32> ret) {
  const auto dims = rew.dimensions();
  if (dims.size() != 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "lrx_gae: rewards must be 2-D");
  const int32_t T = layout == LRX_LAYOUT_TN ? dims[0] : dims[1], N = layout == LRX_LAYOUT_TN ? dims[1] : dims[0];
  return status(lrx_gae(rew.typed_data(), val.typed_data(), done.typed_data(), last_value.typed_data(), adv->typed_data(),
                        ret->typed_data(), N, T, gamma, lam, layout, /*ev_sums=*/nullptr, stream),
                "lrx_gae");
}

// --------------------------------------------------------------------------------------------- lrx_rollout
// algorithm/ppo.py:199-316 under the vmap of :360-368: T steps of PPO.step for N envs + last_value.
ffi::Error RolloutImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Buffer<ffi::F32> params,
                       ffi::Buffer<ffi::F32> y, ffi::Buffer<ffi::F32> t, ffi::Buffer<ffi::S32> step_count,
                       int32_t env_kind, int32_t max_episode_steps, float dt, ffi::Span<const float> env_params,
                       int32_t obs_dim, int32_t n_out, bool is_box, ffi::Span<const int32_t> n_layers,
                       ffi::Span<const int32_t> dims, int64_t seed, int32_t env_offset, int32_t step_offset, int32_t T,
                       float gamma, ffi::ResultBuffer<ffi::F32> y_out, ffi::ResultBuffer<ffi::F32> t_out,
                       ffi::ResultBuffer<ffi::S32> sc_out, ffi::ResultBuffer<ffi::F32> obs, ffi::ResultBuffer<ffi::F32> act,
                       ffi::ResultBuffer<ffi::F32> rew, ffi::ResultBuffer<ffi::U8> done, ffi::ResultBuffer<ffi::F32> logp,
                       ffi::ResultBuffer<ffi::F32> val, ffi::ResultBuffer<ffi::F32> last_value) {
  const int32_t N = static_cast<int32_t>(t.element_count());
  const LrxEnvDesc env = make_env(env_kind, max_episode_steps, dt, env_params);
  const LrxPolicyDesc pol = make_policy(obs_dim, n_out, is_box, n_layers, dims);
  // the carried env state is updated in place by the library: XLA gives distinct result buffers, so copy first
  cudaMemcpyAsync(y_out->typed_data(), y.typed_data(), y.size_bytes(), cudaMemcpyDeviceToDevice, stream);
  cudaMemcpyAsync(t_out->typed_data(), t.typed_data(), t.size_bytes(), cudaMemcpyDeviceToDevice, stream);
  cudaMemcpyAsync(sc_out->typed_data(), step_count.typed_data(), step_count.size_bytes(), cudaMemcpyDeviceToDevice, stream);
  LrxEnvState st{y_out->typed_data(), t_out->typed_data(), sc_out->typed_data()};
  LrxRolloutOut out{};
  out.obs = obs->typed_data();
  out.act = act->untyped_data();  // int32 bits for Discrete actions, float for Box
  out.rew = rew->typed_data();
  out.done = done->typed_data();
  out.logp = logp->typed_data();
  out.val = val->typed_data();
  out.last_value = last_value->typed_data();
  const size_t ws = lrx_rollout_workspace_bytes(&pol, N);
  if (ws) {
    auto mem = scratch.Allocate(ws);
    if (mem.has_value()) { out.workspace = *mem; out.workspace_bytes = ws; }
  }
  const LrxRng rng{static_cast<uint64_t>(seed), static_cast<uint32_t>(env_offset), static_cast<uint32_t>(step_offset)};
  return status(lrx_rollout(&env, &pol, params.typed_data(), st, rng, /*replay=*/nullptr, out, N, T, gamma, stream),
                "lrx_rollout");
}

// ------------------------------------------------------------------------------------------ lrx_ppo_update
// One epoch of train_epoch (algorithm/ppo.py:494-521): for each row of idx, ppo_loss + grad + optax clip + Adam.
ffi::Error PpoUpdateImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Buffer<ffi::F32> params,
                         ffi::Buffer<ffi::F32> mu, ffi::Buffer<ffi::F32> nu, ffi::Buffer<ffi::S32> count,
                         ffi::Buffer<ffi::F32> obs, ffi::Buffer<ffi::F32> act, ffi::Buffer<ffi::F32> logp,
                         ffi::Buffer<ffi::F32> val, ffi::Buffer<ffi::F32> ret, ffi::Buffer<ffi::F32> adv,
                         ffi::Buffer<ffi::S32> idx, int32_t obs_dim, int32_t n_out, bool is_box,
                         ffi::Span<const int32_t> n_layers, ffi::Span<const int32_t> dims, ffi::Span<const float> hyper,
                         bool clip_value_loss, bool normalize_advantages, int32_t surrogate,
                         ffi::ResultBuffer<ffi::F32> params_out, ffi::ResultBuffer<ffi::F32> mu_out,
                         ffi::ResultBuffer<ffi::F32> nu_out, ffi::ResultBuffer<ffi::S32> count_out,
                         ffi::ResultBuffer<ffi::F32> stats, ffi::ResultBuffer<ffi::S32> nonfinite) {
  const LrxPolicyDesc pol = make_policy(obs_dim, n_out, is_box, n_layers, dims);
  const auto od = logp.dimensions();  // [T][N]
  const auto id = idx.dimensions();   // [num_batches][B]
  LrxBatchView view{};
  view.obs = obs.typed_data();
  view.act = act.untyped_data();
  view.logp = logp.typed_data();
  view.val = val.typed_data();
  view.ret = ret.typed_data();
  view.adv = adv.typed_data();
  view.T = static_cast<int32_t>(od[0]);
  view.N = static_cast<int32_t>(od[1]);
  const int32_t num_batches = static_cast<int32_t>(id[0]);
  const int64_t B = id[1];
  // hyper = clip_coefficient, value_loss_coefficient, entropy_loss_coefficient, max_grad_norm, learning_rate, b1, b2, eps
  LrxPpoHyper h{hyper[0], hyper[1], hyper[2], hyper[3], hyper[4], hyper[5], hyper[6], hyper[7],
                clip_value_loss ? 1 : 0, normalize_advantages ? 1 : 0, surrogate};
  cudaMemcpyAsync(params_out->typed_data(), params.typed_data(), params.size_bytes(), cudaMemcpyDeviceToDevice, stream);
  cudaMemcpyAsync(mu_out->typed_data(), mu.typed_data(), mu.size_bytes(), cudaMemcpyDeviceToDevice, stream);
  cudaMemcpyAsync(nu_out->typed_data(), nu.typed_data(), nu.size_bytes(), cudaMemcpyDeviceToDevice, stream);
  cudaMemcpyAsync(count_out->typed_data(), count.typed_data(), count.size_bytes(), cudaMemcpyDeviceToDevice, stream);
  cudaMemsetAsync(nonfinite->typed_data(), 0, sizeof(int32_t), stream);
  const size_t ws_bytes = lrx_ppo_workspace_bytes(&pol, B);
  auto ws = scratch.Allocate(ws_bytes);
  if (!ws.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "lrx_ppo_update: scratch allocation failed");
  return status(lrx_ppo_update(&pol, params_out->typed_data(), mu_out->typed_data(), nu_out->typed_data(),
                               count_out->typed_data(), &view, idx.typed_data(), num_batches, B, &h, stats->typed_data(),
                               nonfinite->typed_data(), *ws, ws_bytes, stream),
                "lrx_ppo_update");
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(LrxGae, GaeImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()   // rewards
                                  .Arg<ffi::Buffer<ffi::F32>>()   // values
                                  .Arg<ffi::Buffer<ffi::U8>>()    // dones
                                  .Arg<ffi::Buffer<ffi::F32>>()   // last_value
                                  .Attr<float>("gamma")
                                  .Attr<float>("lam")
                                  .Attr<int32_t>("layout")
                                  .Ret<ffi::Buffer<ffi::F32>>()   // advantages
                                  .Ret<ffi::Buffer<ffi::F32>>()); // returns

XLA_FFI_DEFINE_HANDLER_SYMBOL(LrxRollout, RolloutImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::Buffer<ffi::F32>>()   // params (flat, Equinox leaf order)
                                  .Arg<ffi::Buffer<ffi::F32>>()   // env state y [k][N]
                                  .Arg<ffi::Buffer<ffi::F32>>()   // t [N]
                                  .Arg<ffi::Buffer<ffi::S32>>()   // step_count [N]
                                  .Attr<int32_t>("env_kind")
                                  .Attr<int32_t>("max_episode_steps")
                                  .Attr<float>("dt")
                                  .Attr<ffi::Span<const float>>("env_params")
                                  .Attr<int32_t>("obs_dim")
                                  .Attr<int32_t>("n_out")
                                  .Attr<bool>("is_box")
                                  .Attr<ffi::Span<const int32_t>>("n_layers")
                                  .Attr<ffi::Span<const int32_t>>("dims")
                                  .Attr<int64_t>("seed")
                                  .Attr<int32_t>("env_offset")
                                  .Attr<int32_t>("step_offset")
                                  .Attr<int32_t>("num_steps")
                                  .Attr<float>("gamma")
                                  .Ret<ffi::Buffer<ffi::F32>>()   // y'
                                  .Ret<ffi::Buffer<ffi::F32>>()   // t'
                                  .Ret<ffi::Buffer<ffi::S32>>()   // step_count'
                                  .Ret<ffi::Buffer<ffi::F32>>()   // observations [T][N][d]
                                  .Ret<ffi::Buffer<ffi::F32>>()   // actions [T][N] (int32 bits for Discrete)
                                  .Ret<ffi::Buffer<ffi::F32>>()   // rewards
                                  .Ret<ffi::Buffer<ffi::U8>>()    // dones
                                  .Ret<ffi::Buffer<ffi::F32>>()   // log_probs
                                  .Ret<ffi::Buffer<ffi::F32>>()   // values
                                  .Ret<ffi::Buffer<ffi::F32>>()); // last_value [N]

XLA_FFI_DEFINE_HANDLER_SYMBOL(LrxPpoUpdate, PpoUpdateImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::Buffer<ffi::F32>>()   // params
                                  .Arg<ffi::Buffer<ffi::F32>>()   // adam mu
                                  .Arg<ffi::Buffer<ffi::F32>>()   // adam nu
                                  .Arg<ffi::Buffer<ffi::S32>>()   // adam count
                                  .Arg<ffi::Buffer<ffi::F32>>()   // observations [T][N][d]
                                  .Arg<ffi::Buffer<ffi::F32>>()   // actions [T][N]
                                  .Arg<ffi::Buffer<ffi::F32>>()   // log_probs
                                  .Arg<ffi::Buffer<ffi::F32>>()   // values
                                  .Arg<ffi::Buffer<ffi::F32>>()   // returns
                                  .Arg<ffi::Buffer<ffi::F32>>()   // advantages
                                  .Arg<ffi::Buffer<ffi::S32>>()   // minibatch indices [num_batches][B], i = n*T + t
                                  .Attr<int32_t>("obs_dim")
                                  .Attr<int32_t>("n_out")
                                  .Attr<bool>("is_box")
                                  .Attr<ffi::Span<const int32_t>>("n_layers")
                                  .Attr<ffi::Span<const int32_t>>("dims")
                                  .Attr<ffi::Span<const float>>("hyper")
                                  .Attr<bool>("clip_value_loss")
                                  .Attr<bool>("normalize_advantages")
                                  .Attr<int32_t>("surrogate")
                                  .Ret<ffi::Buffer<ffi::F32>>()   // params'
                                  .Ret<ffi::Buffer<ffi::F32>>()   // mu'
                                  .Ret<ffi::Buffer<ffi::F32>>()   // nu'
                                  .Ret<ffi::Buffer<ffi::S32>>()   // count'
                                  .Ret<ffi::Buffer<ffi::F32>>()   // stats [num_batches][LRX_PPO_NSTATS]
                                  .Ret<ffi::Buffer<ffi::S32>>()); // non-finite flag (eqx.error_if of ppo.py:413-417)

#endif  // LRX_WITH_XLA_FFI
