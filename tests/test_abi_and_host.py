"""CPU-side checks: the C-ABI library builds/loads and exports every symbol declared in
include/lerax_b200.h; host-side mirror logic (keys, spaces, PPO bookkeeping, descriptors)."""

import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "lerax_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lrx_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    fns = header_functions()
    for must in ("lrx_rollout", "lrx_gae", "lrx_ppo_update", "lrx_ppo_grad", "lrx_ppo_apply", "lrx_last_error"):
        assert must in fns


def test_library_exports_every_declared_symbol():
    from lerax_b200 import _lib as L

    raw = ctypes.CDLL(L.LIB_PATH)
    for name in header_functions():
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
        assert name in L.PROTOTYPES, f"{name} has no ctypes prototype"
    assert sorted(L.PROTOTYPES) == header_functions()
    assert L.lib.lrx_version() == 100


def test_debug_header_symbols_exported_and_separate():
    """include/lerax_b200_debug.h: the toggles come from the product library, the tensor-core self-tests from the separate
    liblerax_b200_test.so -- and the product library does not carry them."""
    from lerax_b200 import _lib as L

    src = open(os.path.join(ROOT, "include", "lerax_b200_debug.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    declared = sorted(set(re.findall(r"\b(lrx_debug_[a-z_0-9]+)\s*\(", src)))
    assert declared == sorted(list(L.DEBUG_HOOKS) + list(L.DEBUG_TESTS))
    raw = ctypes.CDLL(L.LIB_PATH)
    for name in L.DEBUG_HOOKS:
        assert hasattr(raw, name)
    for name in L.DEBUG_TESTS:
        assert not hasattr(raw, name), f"{name} must not ship in the product library"
    dbg = L.debug_lib()
    for name in declared:
        assert hasattr(dbg, name)


def test_struct_sizes_match_header_layout():
    from lerax_b200 import _lib as L

    assert ctypes.sizeof(L.LrxEnvDesc) == 4 + 4 + 4 + 16 * 4
    assert ctypes.sizeof(L.LrxPolicyDesc) == 4 * 4 + 3 * 4 + 3 * 9 * 4 + 3 * 8 * 4
    assert ctypes.sizeof(L.LrxRng) == 16
    assert ctypes.sizeof(L.LrxRolloutOut) == 12 * 8
    assert ctypes.sizeof(L.LrxPpoHyper) == 8 * 4 + 3 * 4
    assert ctypes.sizeof(L.LrxBatchView) == 6 * 8 + 8 + 8


def test_argument_validation_without_gpu():
    """Validation happens before any CUDA call, so it is checkable on the CPU box."""
    from lerax_b200 import _lib as L

    assert L.lib.lrx_policy_num_params(None) == -1
    d = L.LrxPolicyDesc()
    assert L.lib.lrx_policy_forward(d, None, None, None, None, 1, None) < 0
    assert L.lib.lrx_last_error() != b""
    e = L.LrxEnvDesc()
    e.kind = 9
    assert L.lib.lrx_env_reset(e, L.LrxEnvState(), 4, L.LrxRng(), None, None) == -2
    assert b"unsupported env kind" in L.lib.lrx_last_error()
    assert L.lib.lrx_gae(None, None, None, None, None, None, -1, 4, 0.99, 0.95, 0, None, None) == -1


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lerax_b200 import _lib as L
    from lerax_b200.env.classic_control import CartPole

    with pytest.raises(L.LeraxB200Error, match="no CPU fallback"):
        CartPole().initial(key=0)


def test_policy_descriptor_matches_oracle_layout():
    from lerax_b200 import _lib as L
    from lerax_b200.env.classic_control import Acrobot, CartPole, Pendulum
    from lerax_b200.policy import MLPActorCriticPolicy
    from oracle import policy as OP

    for env, n_out, box in ((CartPole(), 2, False), (Pendulum(), 1, True), (Acrobot(), 3, False)):
        for kw in ({}, dict(feature_width=256, value_width=256, action_width=256), dict(action_depth=1, value_depth=3)):
            pol = MLPActorCriticPolicy(env, key=0, device="cpu", **kw)
            spec = OP.make_spec(env.observation_space.flat_size, n_out, box, **kw)
            assert pol.n_params == spec.n_params == L.lib.lrx_policy_num_params(pol.desc())
            offs, ls = spec.offsets()
            assert [(c, i, w, b, k, n) for (c, i, w, b, k, n) in offs] == pol._layout
            assert ls == pol.log_std_offset
            d = pol.desc()
            for c in range(3):
                assert tuple(d.dims[c][: d.n_layers[c] + 1]) == spec.dims[c]
                assert tuple(bool(x) for x in d.relu[c][: d.n_layers[c]]) == spec.relu[c]
            lim = 1 / np.sqrt(env.observation_space.flat_size)
            w0 = pol.encoder.layers[0].weight.numpy()
            assert w0.shape == (spec.dims[0][1], spec.dims[0][0]) and np.abs(w0).max() <= lim


def test_env_descriptors_and_defaults():
    from lerax_b200.env.classic_control import Acrobot, CartPole, ContinuousMountainCar, MountainCar, Pendulum
    from lerax_b200.wrapper import TimeLimit
    from oracle import envs as OE

    for env, o in ((CartPole(), OE.cartpole()), (Pendulum(), OE.pendulum()), (Acrobot(), OE.acrobot()),
                   (MountainCar(), OE.mountain_car()), (ContinuousMountainCar(), OE.continuous_mountain_car())):
        d = env.desc()
        assert d.kind == o.kind and d.dt == np.float32(o.dt) and d.max_episode_steps == 0
        vals = [np.float32(v) for v in o.params.values()]
        assert [d.p[i] for i in range(len(vals))] == vals
    e = TimeLimit(Pendulum(g=5.0), 200)
    assert e.desc().max_episode_steps == 200 and e.desc().p[2] == 5.0 and e.g == 5.0
    with pytest.raises(NotImplementedError):
        CartPole(solver="euler")


def test_ppo_bookkeeping_matches_reference_formulas():
    from lerax_b200.algorithm import PPO

    a = PPO()
    assert (a.num_envs, a.num_steps, a.num_epochs, a.batch_size) == (4, 2048, 10, 256)
    assert a.num_iterations(2**16) == 8 and a.minibatches_per_epoch == 32
    b = PPO(num_envs=5, num_steps=20, num_batches=32)  # 100 // 32 = 3 -> 33 minibatches (base_buffer.py:85-88)
    assert b.batch_size == 3 and b.minibatches_per_epoch == 33
    with pytest.raises(ValueError):
        PPO(num_envs=1, num_steps=4, num_batches=32)
    with pytest.raises(TypeError):
        a.consolidate_callbacks(object())


def test_callbacks_with_custom_on_step_are_rejected():
    from lerax_b200.algorithm import PPO
    from lerax_b200.callback import AbstractCallback, CallbackList, EmptyCallback

    class Bad(AbstractCallback):
        def on_step(self, ctx, *, key=None):
            return 1

    a = PPO()
    assert isinstance(a.consolidate_callbacks(None), CallbackList)
    assert isinstance(a.consolidate_callbacks([EmptyCallback()]), CallbackList)
    with pytest.raises(NotImplementedError):
        a.consolidate_callbacks([Bad()])


def test_keys_split_deterministic_and_distinct():
    from lerax_b200 import random as jr

    k = jr.key(0)
    a, b = jr.split(k, 2)
    assert a != b and jr.split(k, 2) == [a, b]
    assert len({int(x) for x in jr.split(k, 1000)}) == 1000
    assert jr.key(0) != jr.key(1)


def test_spaces():
    from lerax_b200.space import Box, Discrete

    b = Box(-2.0, 2.0)
    assert b.shape == () and b.flat_size == 1 and b.contains(1.0) and not b.contains(3.0)
    assert Box(-np.ones(3), np.ones(3)).flat_size == 3
    d = Discrete(3)
    assert d.n == 3 and d.contains(2) and not d.contains(3)


def test_empty_callback_preserves_empty_state():
    """tests/test_empty_callback.py of the reference."""
    from types import SimpleNamespace

    from lerax_b200.callback import EmptyCallback, EmptyCallbackState, EmptyCallbackStepState, ResetContext

    callback = EmptyCallback()
    callback_state = callback.reset(ResetContext(locals={}), key=0)
    callback_step_state = callback.step_reset(ResetContext(locals={}), key=0)
    step_context = SimpleNamespace(state=callback_step_state)
    iteration_context = SimpleNamespace(state=callback_state)
    training_context = SimpleNamespace(state=callback_state)
    assert isinstance(callback_state, EmptyCallbackState)
    assert isinstance(callback_step_state, EmptyCallbackStepState)
    assert callback.on_step(step_context, key=0) is callback_step_state
    assert callback.on_iteration(iteration_context, key=0) is callback_state
    assert callback.on_training_start(training_context, key=0) is callback_state
    assert callback.on_training_end(training_context, key=0) is callback_state
    assert callback.continue_training(iteration_context, key=0)


def test_schedule_functions():
    """tests/test_curriculum.py:75-91 of the reference."""
    from lerax_b200.curriculum import cosine_schedule, linear_schedule, step_schedule

    lin = linear_schedule(start=0.0, end=1.0, total=100)
    assert np.isclose(lin(0), 0.0) and np.isclose(lin(50), 0.5) and np.isclose(lin(100), 1.0)
    assert np.isclose(lin(200), 1.0)  # clamped
    stp = step_schedule(values=[1.0, 2.0, 3.0], boundaries=[10, 20])
    assert [stp(i) for i in (0, 9, 10, 19, 20)] == [1.0, 1.0, 2.0, 2.0, 3.0]
    cos = cosine_schedule(start=2.0, end=0.0, total=10)
    assert np.isclose(cos(0), 2.0) and np.isclose(cos(5), 1.0, atol=1e-6) and np.isclose(cos(10), 0.0, atol=1e-6)
    with pytest.raises(ValueError):
        step_schedule(values=[1.0], boundaries=[1, 2])


def test_tensorboard_event_file_format(tmp_path):
    """The dependency-free TensorBoardBackend: TFRecord framing with valid masked CRC-32Cs around Event protos that
    decode back to the logged tags / values / steps."""
    import struct

    from lerax_b200.callback.tensorboard import TensorBoardBackend, crc32c, masked_crc32c

    assert crc32c(b"123456789") == 0xE3069283  # the CRC-32C check value
    b = TensorBoardBackend(log_dir=tmp_path)
    b.open("run")
    b.log_hparams({"algorithm.num_envs": 4})
    b.log_scalars({"episode/return": 21.5, "train/value_loss": 0.25}, step=8192)
    b.close()
    files = list((tmp_path / "run").iterdir())
    assert len(files) == 1 and files[0].name.startswith("events.out.tfevents.")
    data = files[0].read_bytes()

    def varint(buf, i):
        n = shift = 0
        while True:
            n |= (buf[i] & 0x7F) << shift
            shift += 7
            i += 1
            if not buf[i - 1] & 0x80:
                return n, i

    def fields(buf):
        i, out = 0, []
        while i < len(buf):
            k, i = varint(buf, i)
            f, w = k >> 3, k & 7
            if w == 0:
                v, i = varint(buf, i)
            elif w == 1:
                v, i = struct.unpack("<d", buf[i:i + 8])[0], i + 8
            elif w == 5:
                v, i = struct.unpack("<f", buf[i:i + 4])[0], i + 4
            else:
                n, i = varint(buf, i)
                v, i = buf[i:i + n], i + n
            out.append((f, v))
        return out

    events, i = [], 0
    while i < len(data):
        (n,) = struct.unpack("<Q", data[i:i + 8])
        assert struct.unpack("<I", data[i + 8:i + 12])[0] == masked_crc32c(data[i:i + 8])
        payload = data[i + 12:i + 12 + n]
        assert struct.unpack("<I", data[i + 12 + n:i + 16 + n])[0] == masked_crc32c(payload)
        events.append(dict(fields(payload)))
        i += 16 + n
    assert events[0][3] == b"brain.Event:2" and len(events) == 4
    scalars = {}
    for ev in events[2:]:
        assert ev[2] == 8192
        val = dict(fields(dict(fields(ev[5]))[1]))
        scalars[val[1].decode()] = val[2]
    assert scalars == {"episode/return": 21.5, "train/value_loss": 0.25}
    text = dict(fields(dict(fields(events[1][5]))[1]))
    assert text[1] == b"hparams" and b"algorithm.num_envs" in dict(fields(text[8]))[8]
