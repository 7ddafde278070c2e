"""CartPole, Pendulum, Acrobot, MountainCar and ContinuousMountainCar with the reference's constructor kwargs and functional
env interface (initial / observation / transition / reward / terminal / truncate), batched
over N envs on the GPU.

Reference: env/classic_control/base_classic_control.py:49-75, cartpole.py:110-203,
pendulum.py:51-126, acrobot.py:114-277, mountain_car.py:106-186, continuous_mountain_car.py:91-174,
env/base_env.py:32-131.  The dynamics run in the
kernels of lerax_b200/csrc/lrx_env.cuh through the C ABI (lrx_env_reset / lrx_env_step /
lrx_env_observe); physical parameters travel in LrxEnvDesc at call time, so mutating an env
attribute between iterations (curriculum) takes effect without recompiling anything.
Only the reference defaults `solver=Tsit5()` with `ConstantStepSize()` are supported.
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from ... import _lib as L
from ...random import Key
from ...space import Box, Discrete


@dataclass
class ClassicControlState:
    """Batched CartPoleState/PendulumState/AcrobotState (+ TimeLimitState.step_count):
    y [k, N] float32, t [N] float32, step_count [N] int32, all CUDA tensors."""

    y: object
    t: object
    step_count: object

    @property
    def num_envs(self) -> int:
        return int(self.t.shape[0])

    def c_state(self) -> L.LrxEnvState:
        return L.LrxEnvState(L.ptr(self.y), L.ptr(self.t), L.ptr(self.step_count))

    def clone(self) -> "ClassicControlState":
        return ClassicControlState(self.y.clone(), self.t.clone(), self.step_count.clone())


class AbstractClassicControlEnv:
    name: str = ""
    kind: int = -1
    state_dim: int = 0
    max_episode_steps: int = 0  # set by wrapper.TimeLimit

    def __init__(self, solver=None, stepsize_controller=None):
        if solver is not None or stepsize_controller is not None:
            raise NotImplementedError(
                "lerax_b200 implements the reference default only: one diffrax.Tsit5 step with ConstantStepSize."
            )

    # ---- descriptor handed to the kernels
    def _params(self) -> list[float]:
        raise NotImplementedError

    def desc(self) -> L.LrxEnvDesc:
        d = L.LrxEnvDesc()
        d.kind = self.kind
        d.max_episode_steps = int(self.max_episode_steps)
        d.dt = float(np.float32(self.dt))
        p = [float(np.float32(x)) for x in self._params()]
        for i in range(L.LRX_ENV_MAX_PARAMS):
            d.p[i] = p[i] if i < len(p) else 0.0
        return d

    # ---- functional interface (env/base_env.py:32-131), batched
    def initial(self, *, key: Key | int, num_envs: int = 1) -> ClassicControlState:
        torch = L.require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device())
        st = ClassicControlState(
            torch.empty((self.state_dim, num_envs), dtype=torch.float32, device=dev),
            torch.empty((num_envs,), dtype=torch.float32, device=dev),
            torch.empty((num_envs,), dtype=torch.int32, device=dev),
        )
        rng = L.LrxRng(int(key) & ((1 << 64) - 1), 0, 0)
        d = self.desc()
        L.check(L.lib.lrx_env_reset(d, st.c_state(), num_envs, rng, None, L.stream_ptr()), "lrx_env_reset")
        return st

    def initial_from_uniform(self, uniform) -> ClassicControlState:
        """env.initial with the uniform draws injected: uniform [N, k] in [0, 1)."""
        torch = L.require_cuda()
        uniform = uniform.to(dtype=torch.float32).contiguous()
        n = uniform.shape[0]
        st = ClassicControlState(
            torch.empty((self.state_dim, n), dtype=torch.float32, device=uniform.device),
            torch.empty((n,), dtype=torch.float32, device=uniform.device),
            torch.empty((n,), dtype=torch.int32, device=uniform.device),
        )
        d = self.desc()
        L.check(L.lib.lrx_env_reset(d, st.c_state(), n, L.LrxRng(0, 0, 0), L.ptr(uniform), L.stream_ptr()),
                "lrx_env_reset")
        return st

    def observation(self, state: ClassicControlState, *, key=None):
        torch = L.require_cuda()
        n = state.num_envs
        obs = torch.empty((n, self.observation_space.flat_size), dtype=torch.float32, device=state.y.device)
        d = self.desc()
        L.check(L.lib.lrx_env_observe(d, state.c_state(), L.ptr(obs), None, n, L.stream_ptr()), "lrx_env_observe")
        return obs

    def step_full(self, state: ClassicControlState, action):
        """transition + reward + terminal + truncate + next observation in one launch.
        Returns (next_state, reward, terminal, truncated, next_obs); `state` is not modified."""
        torch = L.require_cuda()
        n = state.num_envs
        nxt = state.clone()
        dev = state.y.device
        act_dtype = torch.float32 if isinstance(self.action_space, Box) else torch.int32
        action = action.to(device=dev, dtype=act_dtype).reshape(n).contiguous()
        reward = torch.empty((n,), dtype=torch.float32, device=dev)
        term = torch.empty((n,), dtype=torch.uint8, device=dev)
        trunc = torch.empty((n,), dtype=torch.uint8, device=dev)
        obs = torch.empty((n, self.observation_space.flat_size), dtype=torch.float32, device=dev)
        d = self.desc()
        L.check(L.lib.lrx_env_step(d, nxt.c_state(), L.ptr(action), L.ptr(reward), L.ptr(term), L.ptr(trunc),
                                   L.ptr(obs), n, L.stream_ptr()), "lrx_env_step")
        return nxt, reward, term.bool(), trunc.bool(), obs

    def transition(self, state: ClassicControlState, action, *, key=None) -> ClassicControlState:
        return self.step_full(state, action)[0]

    def reward(self, state, action, next_state, *, key=None):
        return self.step_full(state, action)[1]

    def terminal(self, state: ClassicControlState, *, key=None):
        torch = L.require_cuda()
        n = state.num_envs
        term = torch.empty((n,), dtype=torch.uint8, device=state.y.device)
        d = self.desc()
        L.check(L.lib.lrx_env_observe(d, state.c_state(), None, L.ptr(term), n, L.stream_ptr()), "lrx_env_observe")
        return term.bool()

    def truncate(self, state: ClassicControlState):
        torch = L.require_cuda()
        if self.max_episode_steps <= 0:
            return torch.zeros_like(state.step_count, dtype=torch.bool)
        return state.step_count >= self.max_episode_steps

    def action_mask(self, state, *, key=None):
        return None

    @property
    def unwrapped(self):
        return self


class CartPole(AbstractClassicControlEnv):
    name = "CartPole"
    kind = L.LRX_ENV_CARTPOLE
    state_dim = 4

    def __init__(self, *, gravity=9.8, cart_mass=1.0, pole_mass=0.1, half_length=0.5, force_mag=10.0,
                 theta_threshold_radians=12 * 2 * math.pi / 360, x_threshold=2.4, dt=0.02,
                 solver=None, stepsize_controller=None):
        super().__init__(solver, stepsize_controller)
        self.gravity, self.cart_mass, self.pole_mass = gravity, cart_mass, pole_mass
        self.length, self.force_mag = half_length, force_mag
        self.theta_threshold_radians, self.x_threshold, self.dt = theta_threshold_radians, x_threshold, dt
        self.action_space = Discrete(2)
        high = np.array([x_threshold * 2, np.inf, theta_threshold_radians * 2, np.inf], np.float32)
        self.observation_space = Box(-high, high)

    @property
    def total_mass(self):
        return self.pole_mass + self.cart_mass

    @property
    def polemass_length(self):
        return self.pole_mass * self.length

    def _params(self):
        return [self.gravity, self.cart_mass, self.pole_mass, self.length, self.force_mag,
                self.theta_threshold_radians, self.x_threshold]


class Pendulum(AbstractClassicControlEnv):
    name = "Pendulum"
    kind = L.LRX_ENV_PENDULUM
    state_dim = 2

    def __init__(self, *, max_speed=8.0, max_torque=2.0, g=9.8, m=1.0, l=1.0, dt=0.05,  # noqa: E741
                 solver=None, stepsize_controller=None):
        super().__init__(solver, stepsize_controller)
        self.max_speed, self.max_torque, self.g, self.m, self.l, self.dt = max_speed, max_torque, g, m, l, dt
        self.action_space = Box(-max_torque, max_torque)
        high = np.array([1.0, 1.0, max_speed], np.float32)
        self.observation_space = Box(-high, high)

    def _params(self):
        return [self.max_speed, self.max_torque, self.g, self.m, self.l]


class Acrobot(AbstractClassicControlEnv):
    name = "Acrobot"
    kind = L.LRX_ENV_ACROBOT
    state_dim = 4

    def __init__(self, *, gravity=9.8, link_length_1=1.0, link_length_2=1.0, link_mass_1=1.0, link_mass_2=1.0,
                 link_com_pos_1=0.5, link_com_pos_2=0.5, link_moi=1.0, max_vel_1=4 * math.pi,
                 max_vel_2=9 * math.pi, torque_max_noise=0.0, torques=(-1.0, 0.0, 1.0), dt=0.2,
                 solver=None, stepsize_controller=None):
        super().__init__(solver, stepsize_controller)
        self.gravity = gravity
        self.link_length_1, self.link_length_2 = link_length_1, link_length_2
        self.link_mass_1, self.link_mass_2 = link_mass_1, link_mass_2
        self.link_com_pos_1, self.link_com_pos_2 = link_com_pos_1, link_com_pos_2
        self.link_moi, self.max_vel_1, self.max_vel_2 = link_moi, max_vel_1, max_vel_2
        self.torque_max_noise = torque_max_noise  # stored but unused, as in the reference
        self.torques = tuple(float(x) for x in torques)
        if len(self.torques) != 3:
            raise ValueError("Acrobot needs exactly 3 torques")
        self.dt = dt
        self.action_space = Discrete(3)
        high = np.array([1.0, 1.0, 1.0, 1.0, max_vel_1, max_vel_2], np.float32)
        self.observation_space = Box(-high, high)

    def _params(self):
        return [self.gravity, self.link_length_1, self.link_length_2, self.link_mass_1, self.link_mass_2,
                self.link_com_pos_1, self.link_com_pos_2, self.link_moi, self.max_vel_1, self.max_vel_2,
                *self.torques]


class MountainCar(AbstractClassicControlEnv):
    """mountain_car.py:32-186: three discrete actions (push left / none / right), reward -1 per step."""

    name = "MountainCar"
    kind = L.LRX_ENV_MOUNTAINCAR
    state_dim = 2

    def __init__(self, *, min_position=-1.2, max_position=0.6, max_speed=0.07, goal_position=0.5, goal_velocity=0.0,
                 force=0.001, gravity=0.0025, dt=1.0, solver=None, stepsize_controller=None):
        super().__init__(solver, stepsize_controller)
        self.min_position, self.max_position, self.max_speed = min_position, max_position, max_speed
        self.goal_position, self.goal_velocity = goal_position, goal_velocity
        self.force, self.gravity, self.dt = force, gravity, dt
        self.low = np.array([min_position, -max_speed], np.float32)
        self.high = np.array([max_position, max_speed], np.float32)
        self.action_space = Discrete(3)
        self.observation_space = Box(self.low, self.high)

    def _params(self):
        return [self.min_position, self.max_position, self.max_speed, self.goal_position, self.goal_velocity,
                self.force, self.gravity]


class ContinuousMountainCar(AbstractClassicControlEnv):
    """continuous_mountain_car.py:32-174: scalar Box action (engine power), reward -0.1 a^2 (+100 at the goal)."""

    name = "ContinuousMountainCar"
    kind = L.LRX_ENV_CONTINUOUS_MOUNTAINCAR
    state_dim = 2

    def __init__(self, *, min_action=-1.0, max_action=1.0, min_position=-1.2, max_position=0.6, max_speed=0.07,
                 goal_position=0.5, goal_velocity=0.0, power=0.0015, dt=1.0, solver=None, stepsize_controller=None):
        super().__init__(solver, stepsize_controller)
        self.min_action, self.max_action = min_action, max_action
        self.min_position, self.max_position, self.max_speed = min_position, max_position, max_speed
        self.goal_position, self.goal_velocity, self.power, self.dt = goal_position, goal_velocity, power, dt
        self.low = np.array([min_position, -max_speed], np.float32)
        self.high = np.array([max_position, max_speed], np.float32)
        self.action_space = Box(min_action, max_action)
        self.observation_space = Box(self.low, self.high)

    def _params(self):
        return [self.min_action, self.max_action, self.min_position, self.max_position, self.max_speed,
                self.goal_position, self.goal_velocity, self.power]


__all__ = ["Acrobot", "CartPole", "ContinuousMountainCar", "MountainCar", "Pendulum", "ClassicControlState",
           "AbstractClassicControlEnv"]
