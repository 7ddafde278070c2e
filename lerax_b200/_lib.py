"""ctypes binding of include/lerax_b200.h (the C-ABI drop-in boundary).

There is deliberately no fallback: if the shared library is missing this module raises, and
every compute entry point raises if no CUDA device is present.
"""

from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_C", "liblerax_b200.so")

LRX_OK = 0
LRX_ENV_CARTPOLE, LRX_ENV_PENDULUM, LRX_ENV_ACROBOT = 0, 1, 2
LRX_ENV_MOUNTAINCAR, LRX_ENV_CONTINUOUS_MOUNTAINCAR = 3, 4
LRX_ENV_MAX_PARAMS = 16
LRX_MAX_LAYERS = 8
LRX_MAX_NOUT = 8
LRX_LAYOUT_TN, LRX_LAYOUT_NT = 0, 1
LRX_P2P_HANDLE_BYTES = 64
LRX_PPO_NSTATS = 8
LRX_PACKED_FLOATS = 16
LRX_GAE_EV_BLOCKS = 2048
LRX_GAE_EV_DOUBLES = 4 + 4 * LRX_GAE_EV_BLOCKS + 4


class LrxEnvDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("max_episode_steps", C.c_int32), ("dt", C.c_float),
                ("p", C.c_float * LRX_ENV_MAX_PARAMS)]


class LrxEnvState(C.Structure):
    _fields_ = [("y", C.c_void_p), ("t", C.c_void_p), ("step_count", C.c_void_p)]


class LrxPolicyDesc(C.Structure):
    _fields_ = [("obs_dim", C.c_int32), ("n_out", C.c_int32), ("is_box", C.c_int32), ("n_params", C.c_int32),
                ("n_layers", C.c_int32 * 3),
                ("dims", (C.c_int32 * (LRX_MAX_LAYERS + 1)) * 3),
                ("relu", (C.c_int32 * LRX_MAX_LAYERS) * 3)]


class LrxRng(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("env_offset", C.c_uint32), ("step_offset", C.c_uint32)]


class LrxReplay(C.Structure):
    _fields_ = [("action_noise", C.c_void_p), ("reset_uniform", C.c_void_p)]


class LrxRolloutOut(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("act", C.c_void_p), ("rew", C.c_void_p), ("done", C.c_void_p),
                ("logp", C.c_void_p), ("val", C.c_void_p), ("last_value", C.c_void_p),
                ("y_trace", C.c_void_p), ("t_trace", C.c_void_p), ("y_next_trace", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class LrxPpoHyper(C.Structure):
    _fields_ = [("clip_coefficient", C.c_float), ("value_loss_coefficient", C.c_float),
                ("entropy_loss_coefficient", C.c_float), ("max_grad_norm", C.c_float),
                ("learning_rate", C.c_float), ("adam_b1", C.c_float), ("adam_b2", C.c_float),
                ("adam_eps", C.c_float), ("clip_value_loss", C.c_int32), ("normalize_advantages", C.c_int32),
                ("surrogate", C.c_int32)]


class LrxEpisodeStats(C.Structure):
    _fields_ = [("step", C.c_void_p), ("episode_return", C.c_void_p), ("episode_length", C.c_void_p),
                ("episode_done", C.c_void_p), ("average_return", C.c_void_p), ("average_length", C.c_void_p)]


class LrxBatchView(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("act", C.c_void_p), ("logp", C.c_void_p), ("val", C.c_void_p),
                ("ret", C.c_void_p), ("adv", C.c_void_p), ("N", C.c_int32), ("T", C.c_int32), ("packed", C.c_void_p)]


_P = C.POINTER
_vp = C.c_void_p

# name -> (restype, argtypes); must list every function declared in include/lerax_b200.h
PROTOTYPES = {
    "lrx_policy_num_params": (C.c_int, [_P(LrxPolicyDesc)]),
    "lrx_env_reset": (C.c_int, [_P(LrxEnvDesc), LrxEnvState, C.c_int32, LrxRng, _vp, _vp]),
    "lrx_env_step": (C.c_int, [_P(LrxEnvDesc), LrxEnvState, _vp, _vp, _vp, _vp, _vp, C.c_int32, _vp]),
    "lrx_env_observe": (C.c_int, [_P(LrxEnvDesc), LrxEnvState, _vp, _vp, C.c_int32, _vp]),
    "lrx_policy_forward": (C.c_int, [_P(LrxPolicyDesc), _vp, _vp, _vp, _vp, C.c_int32, _vp]),
    "lrx_rollout": (C.c_int, [_P(LrxEnvDesc), _P(LrxPolicyDesc), _vp, LrxEnvState, LrxRng, _P(LrxReplay),
                              LrxRolloutOut, C.c_int32, C.c_int32, C.c_float, _vp]),
    "lrx_rollout_workspace_bytes": (C.c_size_t, [_P(LrxPolicyDesc), C.c_int32]),
    "lrx_gae": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32,
                          _vp, _vp]),
    "lrx_returns": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_int32, _vp]),
    "lrx_permutation": (C.c_int, [_vp, C.c_int64, C.c_uint64, _vp]),
    "lrx_p2p_buffer_bytes": (C.c_size_t, [C.c_int32]),
    "lrx_p2p_alloc": (C.c_int, [C.c_int32, _P(C.c_void_p), _vp]),
    "lrx_p2p_open": (C.c_int, [_vp, _P(C.c_void_p)]),
    "lrx_p2p_close": (C.c_int, [_vp]),
    "lrx_p2p_free": (C.c_int, [_vp]),
    "lrx_p2p_slot": (C.c_void_p, [_vp, C.c_int32, C.c_uint32]),
    "lrx_p2p_allreduce": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int32, _vp, C.c_uint32, _vp, _vp]),
    "lrx_ppo_grad_p2p": (C.c_int, [_P(LrxPolicyDesc), _vp, _P(LrxBatchView), _vp, C.c_int64, C.c_int64, _vp, _P(LrxPpoHyper), _vp,
                                    C.c_uint32, _vp, C.c_size_t, _vp]),
    "lrx_ppo_update_p2p": (C.c_int, [_P(LrxPolicyDesc), _vp, _vp, _vp, _vp, _P(LrxBatchView), _vp, C.c_int32, C.c_int64, C.c_int64,
                                      _vp, _P(LrxPpoHyper), _vp, _vp, _vp, C.c_int32, C.c_int32, C.c_uint64, _vp, _vp, _vp, _vp,
                                      C.c_size_t, _vp, C.c_size_t, _vp]),
    "lrx_ppo_grad_apply_p2p": (C.c_int, [_P(LrxPolicyDesc), _vp, _vp, _vp, _vp, _P(LrxBatchView), _vp, C.c_int64, C.c_int64, _vp,
                                          _P(LrxPpoHyper), _vp, _vp, C.c_int32, C.c_int32, C.c_uint32, _vp, _vp, _vp, _vp,
                                          C.c_size_t, _vp, _P(C.c_int32)]),
    "lrx_ppo_apply_p2p_scratch_bytes": (C.c_size_t, [_P(LrxPolicyDesc)]),
    "lrx_ppo_apply_p2p": (C.c_int, [_P(LrxPolicyDesc), _vp, _vp, _vp, _vp, _vp, C.c_int32, C.c_int32, C.c_uint32,
                                    _P(LrxPpoHyper), _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "lrx_episode_stats": (C.c_int, [_vp, _vp, _P(LrxEpisodeStats), C.c_int32, C.c_int32, C.c_float, C.c_int32, _vp]),
    "lrx_ppo_workspace_bytes": (C.c_size_t, [_P(LrxPolicyDesc), C.c_int64]),
    "lrx_ppo_pack_rows": (C.c_int, [_P(LrxBatchView), C.c_int32, C.c_int32, _vp, _vp]),
    "lrx_ppo_adv_sums": (C.c_int, [_vp, C.c_int32, C.c_int32, _vp, C.c_int32, C.c_int64, _vp, _vp]),
    "lrx_ppo_grad": (C.c_int, [_P(LrxPolicyDesc), _vp, _P(LrxBatchView), _vp, C.c_int64, C.c_int64, _vp,
                               _P(LrxPpoHyper), _vp, _vp, C.c_size_t, _vp]),
    "lrx_ppo_apply": (C.c_int, [_P(LrxPolicyDesc), _vp, _vp, _vp, _vp, _vp, _P(LrxPpoHyper), _vp, _vp, _vp]),
    "lrx_ppo_update": (C.c_int, [_P(LrxPolicyDesc), _vp, _vp, _vp, _vp, _P(LrxBatchView), _vp, C.c_int32,
                                 C.c_int64, _P(LrxPpoHyper), _vp, _vp, _vp, C.c_size_t, _vp]),
    "lrx_ppo_update_schedule": (C.c_int, [_P(LrxPolicyDesc), _vp, _vp, _vp, _vp, _P(LrxBatchView), _vp, C.c_int32,
                                          C.c_int64, _P(LrxPpoHyper), _P(C.c_float), _vp, _vp, _vp, C.c_size_t, _vp]),
    "lrx_last_error": (C.c_char_p, []),
    "lrx_version": (C.c_int, []),
    "lrx_launch_count": (C.c_int64, [C.c_int]),
}


# include/lerax_b200_debug.h: test hooks exported by the product library ...
DEBUG_HOOKS = {
    "lrx_debug_set_fused": (None, [C.c_int]),
    "lrx_debug_copy_f32": (C.c_int, [_vp, _vp, C.c_int32, _vp]),
    "lrx_debug_gemm": (C.c_int, [_vp, C.c_int32, _vp, C.c_int32, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_int32, _vp, C.c_int32, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                 C.c_int32, _vp]),
    "lrx_debug_set_gae_chunk": (None, [C.c_int]),
    "lrx_debug_set_tc128_prof": (None, [_vp]),
    "lrx_debug_set_pp64_prof": (None, [_vp, C.c_int]),
}
# ... and the tensor-core self-tests of the separate liblerax_b200_test.so
DEBUG_TESTS = {
    "lrx_debug_umma_gemm": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, _P(C.c_int32), C.c_int32,
                                      _vp]),
    "lrx_debug_umma_gemm_bf16": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, _P(C.c_int32),
                                           C.c_int32, _vp]),
    "lrx_debug_umma_time_bf16": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, _P(C.c_int32),
                                           C.c_int32, C.c_int32, _vp, _vp]),
    "lrx_debug_umma_rate": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _vp, _vp]),
    "lrx_debug_tmem_ld_under_mma": (C.c_int, [C.c_int32, _vp, _vp]),
    "lrx_debug_tmem_shape_probe": (C.c_int, [_vp, _vp]),
    "lrx_debug_tmem_operand_probe": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int32, _vp]),
}


class LeraxB200Error(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"lerax_b200: CUDA library not built ({LIB_PATH} missing). Run `python -m lerax_b200.build` "
            "(needs nvcc). There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()
TEST_LIB_PATH = os.path.join(HERE, "_C", "liblerax_b200_test.so")
_debug = None


def debug_lib():
    """Functions of include/lerax_b200_debug.h (tests and tools only): the hooks of the product library plus the
    self-tests of liblerax_b200_test.so, as attributes of one object."""
    global _debug
    if _debug is None:
        class _Debug:
            pass

        d = _Debug()
        for name, (res, args) in DEBUG_HOOKS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
            setattr(d, name, fn)
        if not os.path.exists(TEST_LIB_PATH):
            raise ImportError(f"lerax_b200: {TEST_LIB_PATH} missing; run `python -m lerax_b200.build`")
        tl = C.CDLL(TEST_LIB_PATH)
        for name, (res, args) in DEBUG_TESTS.items():
            fn = getattr(tl, name)
            fn.restype, fn.argtypes = res, args
            setattr(d, name, fn)
        _debug = d
    return _debug


def check(rc: int, what: str = "") -> None:
    if rc != LRX_OK:
        msg = lib.lrx_last_error().decode("utf-8", "replace")
        raise LeraxB200Error(f"{what or 'lerax_b200'} failed (code {rc}): {msg}")


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise LeraxB200Error("lerax_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
    return torch


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "lerax_b200 expects contiguous CUDA tensors"
    return t.data_ptr()


def stream_ptr():
    import torch

    return torch.cuda.current_stream().cuda_stream
