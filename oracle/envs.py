"""Classic-control environments (CartPole, Pendulum, Acrobot, MountainCar, ContinuousMountainCar),
vectorised over envs.

TEST INFRASTRUCTURE (see oracle/__init__.py).  State arrays are ``y[k, N]`` (state
component on axis 0).  Each function cites the reference lines it restates; the
operation ORDER of the reference expressions is kept so fp32 rounding matches as
closely as a non-XLA evaluation can.

    base:     /root/reference/src/lerax/env/classic_control/base_classic_control.py:49-75
    CartPole: /root/reference/src/lerax/env/classic_control/cartpole.py:110-203
    Pendulum: /root/reference/src/lerax/env/classic_control/pendulum.py:51-126
    Acrobot:  /root/reference/src/lerax/env/classic_control/acrobot.py:114-277
    MountainCar: /root/reference/src/lerax/env/classic_control/mountain_car.py:106-186
    ContinuousMountainCar: /root/reference/src/lerax/env/classic_control/continuous_mountain_car.py:91-174
    TimeLimit:/root/reference/src/lerax/wrapper/misc.py:115-206
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import tsit5

CARTPOLE, PENDULUM, ACROBOT, MOUNTAINCAR, CONTINUOUS_MOUNTAINCAR = 0, 1, 2, 3, 4


def _jnp_mod(x, m):
    """jnp.remainder for floats: C fmod, then shifted to take the sign of the divisor."""
    r = np.fmod(x, m)
    fix = (r != 0) & ((r < 0) != (m < 0))
    return np.where(fix, r + m, r)


@dataclass
class EnvSpec:
    """Physical parameters of one environment family (runtime values, never constants:
    curriculum callbacks rewrite them between iterations, curriculum/scheduled.py:133-138)."""

    kind: int
    params: dict
    dt: float
    max_episode_steps: int = 0  # 0 = no TimeLimit wrapper
    state_dim: int = field(init=False)
    obs_dim: int = field(init=False)
    n_actions: int = field(init=False)  # 0 for Box
    is_box: bool = field(init=False)

    def __post_init__(self):
        self.state_dim = {CARTPOLE: 4, PENDULUM: 2, ACROBOT: 4, MOUNTAINCAR: 2, CONTINUOUS_MOUNTAINCAR: 2}[self.kind]
        self.obs_dim = {CARTPOLE: 4, PENDULUM: 3, ACROBOT: 6, MOUNTAINCAR: 2, CONTINUOUS_MOUNTAINCAR: 2}[self.kind]
        self.n_actions = {CARTPOLE: 2, PENDULUM: 0, ACROBOT: 3, MOUNTAINCAR: 3, CONTINUOUS_MOUNTAINCAR: 0}[self.kind]
        self.is_box = self.kind in (PENDULUM, CONTINUOUS_MOUNTAINCAR)

    def p(self, name, dtype):
        return dtype(np.float32(self.params[name]))  # reference stores fp32 jnp scalars


def cartpole(**kw) -> EnvSpec:
    # cartpole.py:110-122 defaults
    params = dict(
        gravity=9.8, cart_mass=1.0, pole_mass=0.1, half_length=0.5, force_mag=10.0,
        theta_threshold_radians=12 * 2 * np.pi / 360, x_threshold=2.4,
    )
    dt = kw.pop("dt", 0.02)
    mes = kw.pop("max_episode_steps", 0)
    params.update(kw)
    return EnvSpec(CARTPOLE, params, dt, mes)


def pendulum(**kw) -> EnvSpec:
    # pendulum.py:51-62 defaults
    params = dict(max_speed=8.0, max_torque=2.0, g=9.8, m=1.0, l=1.0)
    dt = kw.pop("dt", 0.05)
    mes = kw.pop("max_episode_steps", 0)
    params.update(kw)
    return EnvSpec(PENDULUM, params, dt, mes)


def acrobot(**kw) -> EnvSpec:
    # acrobot.py:114-132 defaults
    params = dict(
        gravity=9.8, link_length_1=1.0, link_length_2=1.0, link_mass_1=1.0, link_mass_2=1.0,
        link_com_pos_1=0.5, link_com_pos_2=0.5, link_moi=1.0,
        max_vel_1=4 * np.pi, max_vel_2=9 * np.pi, torque_0=-1.0, torque_1=0.0, torque_2=1.0,
    )
    dt = kw.pop("dt", 0.2)
    mes = kw.pop("max_episode_steps", 0)
    params.update(kw)
    return EnvSpec(ACROBOT, params, dt, mes)


def mountain_car(**kw) -> EnvSpec:
    # mountain_car.py:106-118 defaults
    params = dict(min_position=-1.2, max_position=0.6, max_speed=0.07, goal_position=0.5, goal_velocity=0.0,
                  force=0.001, gravity=0.0025)
    dt = kw.pop("dt", 1.0)
    mes = kw.pop("max_episode_steps", 0)
    params.update(kw)
    return EnvSpec(MOUNTAINCAR, params, dt, mes)


def continuous_mountain_car(**kw) -> EnvSpec:
    # continuous_mountain_car.py:91-104 defaults
    params = dict(min_action=-1.0, max_action=1.0, min_position=-1.2, max_position=0.6, max_speed=0.07,
                  goal_position=0.5, goal_velocity=0.0, power=0.0015)
    dt = kw.pop("dt", 1.0)
    mes = kw.pop("max_episode_steps", 0)
    params.update(kw)
    return EnvSpec(CONTINUOUS_MOUNTAINCAR, params, dt, mes)


# ----------------------------------------------------------------------------- initial
def initial_from_uniform(spec: EnvSpec, u, dtype=np.float32):
    """env.initial with the uniform draws u[k, N] in [0,1) injected.

    jax.random.uniform(key, shape, minval, maxval) = max(minval, u*(maxval-minval)+minval).
    cartpole.py:157-160, pendulum.py:84-87, acrobot.py:162-165.  Also sets t = 0.
    """
    u = np.asarray(u, dtype=dtype)
    if spec.kind == CARTPOLE:
        lo = np.full((4, 1), -0.05, dtype)
        hi = np.full((4, 1), 0.05, dtype)
    elif spec.kind == PENDULUM:
        hi = np.array([[np.pi], [1.0]], dtype)
        lo = -hi
    elif spec.kind in (MOUNTAINCAR, CONTINUOUS_MOUNTAINCAR):
        # mountain_car.py:145-149, continuous_mountain_car.py:132-136: x ~ U(-0.6, -0.4), v = 0 (u[1] unused)
        lo0, hi0 = u.dtype.type(-0.6), u.dtype.type(-0.4)
        x = np.maximum(lo0, u[0] * (hi0 - lo0) + lo0)
        return np.stack([x, np.zeros_like(x)])
    else:
        lo = np.full((4, 1), -0.1, dtype)
        hi = np.full((4, 1), 0.1, dtype)
    return np.maximum(lo, u * (hi - lo) + lo)


# ---------------------------------------------------------------------------- dynamics
def dynamics(spec: EnvSpec, y, action):
    dt = y.dtype.type
    if spec.kind == CARTPOLE:
        # cartpole.py:162-177
        g = spec.p("gravity", dt)
        m_c = spec.p("cart_mass", dt)
        m_p = spec.p("pole_mass", dt)
        length = spec.p("half_length", dt)
        f_mag = spec.p("force_mag", dt)
        total_mass = m_p + m_c
        pml = m_p * length
        _, x_dot, theta, theta_dot = y
        force = (action * 2 - 1).astype(y.dtype) * f_mag
        sin, cos = np.sin(theta), np.cos(theta)
        temp = (force + pml * (theta_dot * theta_dot) * sin) / total_mass
        theta_dd = (g * sin - cos * temp) / (length * (dt(4.0 / 3.0) - m_p * (cos * cos) / total_mass))
        x_dd = temp - pml * theta_dd * cos / total_mass
        return np.stack([x_dot, x_dd, theta_dot, theta_dd])
    if spec.kind == PENDULUM:
        # pendulum.py:89-98
        g, m, l = spec.p("g", dt), spec.p("m", dt), spec.p("l", dt)
        mt = spec.p("max_torque", dt)
        theta, theta_dot = y
        u = np.clip(action.astype(y.dtype), -mt, mt)
        theta_dd = dt(3.0) * g / (dt(2.0) * l) * np.sin(theta) + dt(3.0) / (m * (l * l)) * u
        return np.stack([theta_dot, theta_dd])
    if spec.kind == MOUNTAINCAR:
        # mountain_car.py:152-158
        x, x_d = y
        u = (action - 1).astype(y.dtype) * spec.p("force", dt)
        x_dd = u - spec.p("gravity", dt) * np.cos(dt(3.0) * x)
        return np.stack([x_d, x_dd])
    if spec.kind == CONTINUOUS_MOUNTAINCAR:
        # continuous_mountain_car.py:138-144
        x, x_d = y
        a = np.clip(action.astype(y.dtype), spec.p("min_action", dt), spec.p("max_action", dt))
        x_dd = spec.p("power", dt) * a - dt(0.0025) * np.cos(dt(3.0) * x)
        return np.stack([x_d, x_dd])
    # acrobot.py:167-232
    g = spec.p("gravity", dt)
    l1 = spec.p("link_length_1", dt)
    m1, m2 = spec.p("link_mass_1", dt), spec.p("link_mass_2", dt)
    lc1, lc2 = spec.p("link_com_pos_1", dt), spec.p("link_com_pos_2", dt)
    moi = spec.p("link_moi", dt)
    torques = np.array([spec.p("torque_0", dt), spec.p("torque_1", dt), spec.p("torque_2", dt)], dtype=y.dtype)
    theta1, theta2, theta1_d, theta2_d = y
    a = torques[action]
    two = dt(2.0)
    half_pi = dt(np.pi / 2)
    cos2, sin2 = np.cos(theta2), np.sin(theta2)
    d1 = m1 * (lc1 * lc1) + m2 * ((l1 * l1) + (lc2 * lc2) + two * l1 * lc2 * cos2) + moi + moi
    d2 = m2 * ((lc2 * lc2) + l1 * lc2 * cos2) + moi
    phi2 = m2 * lc2 * g * np.cos(theta1 + theta2 - half_pi)
    phi1 = (
        -m2 * l1 * lc2 * (theta2_d * theta2_d) * sin2
        - two * m2 * l1 * lc2 * theta1_d * theta2_d * sin2
        + (m1 * lc1 + m2 * l1) * g * np.cos(theta1 - half_pi)
        + phi2
    )
    theta2_dd = (a + d2 / d1 * phi1 - m2 * l1 * lc2 * (theta1_d * theta1_d) * sin2 - phi2) / (
        m2 * (lc2 * lc2) + moi - (d2 * d2) / d1
    )
    theta1_dd = -(d2 * theta2_dd + phi1) / d1
    return np.stack([theta1_d, theta2_d, theta1_dd, theta2_dd])


def clip_state(spec: EnvSpec, y):
    dt = y.dtype.type
    if spec.kind == CARTPOLE:
        return y  # cartpole.py:179-180
    pi, two_pi = dt(np.pi), dt(2 * np.pi)
    if spec.kind == PENDULUM:
        # pendulum.py:100-104
        ms = spec.p("max_speed", dt)
        theta = _jnp_mod(y[0] + pi, two_pi) - pi
        return np.stack([theta, np.clip(y[1], -ms, ms)])
    if spec.kind == MOUNTAINCAR:
        # mountain_car.py:160-165
        ms, lo, hi = spec.p("max_speed", dt), spec.p("min_position", dt), spec.p("max_position", dt)
        v = np.clip(y[1], -ms, ms)
        x = np.clip(y[0], lo, hi)
        v = v * ((x != lo) | (v > dt(0.0))).astype(y.dtype)
        return np.stack([x, v])
    if spec.kind == CONTINUOUS_MOUNTAINCAR:
        # continuous_mountain_car.py:146-150
        ms, lo, hi = spec.p("max_speed", dt), spec.p("min_position", dt), spec.p("max_position", dt)
        return np.stack([np.clip(y[0], lo, hi), np.clip(y[1], -ms, ms)])
    # acrobot.py:234-241
    mv1, mv2 = spec.p("max_vel_1", dt), spec.p("max_vel_2", dt)
    a1 = _jnp_mod(y[0] + pi, two_pi) - pi
    a2 = _jnp_mod(y[1] + pi, two_pi) - pi
    return np.stack([a1, a2, np.clip(y[2], -mv1, mv1), np.clip(y[3], -mv2, mv2)])


def transition(spec: EnvSpec, y, t, action):
    """base_classic_control.py:49-72: one Tsit5 step of size h = (t+dt)-t, clip, t += dt."""
    dtc = y.dtype.type(np.float32(spec.dt))
    h = tsit5.step_size(t, dtc)
    y1 = tsit5.step(lambda yy: dynamics(spec, yy, action), y, h)
    return clip_state(spec, y1), t + dtc


def observation(spec: EnvSpec, y):
    if spec.kind == CARTPOLE:
        return y  # cartpole.py:182-185
    if spec.kind == PENDULUM:
        return np.stack([np.cos(y[0]), np.sin(y[0]), y[1]])  # pendulum.py:106-110
    if spec.kind in (MOUNTAINCAR, CONTINUOUS_MOUNTAINCAR):
        return y  # mountain_car.py:167-170, continuous_mountain_car.py:152-155
    # acrobot.py:243-256
    return np.stack([np.cos(y[0]), np.sin(y[0]), np.cos(y[1]), np.sin(y[1]), y[2], y[3]])


def reward(spec: EnvSpec, y, action, y_next):
    dt = y.dtype.type
    if spec.kind == CARTPOLE:
        return np.ones(y.shape[1], dtype=y.dtype)  # cartpole.py:187-195
    if spec.kind == PENDULUM:
        # pendulum.py:112-123 (uses the NEXT state)
        mt = spec.p("max_torque", dt)
        theta, theta_dot = y_next
        u = np.clip(action.astype(y.dtype), -mt, mt)
        cost = theta * theta + dt(0.1) * (theta_dot * theta_dot) + dt(0.001) * (u * u)
        return -cost
    if spec.kind == MOUNTAINCAR:
        return np.full(y.shape[1], -1.0, dtype=y.dtype)  # mountain_car.py:172-180
    if spec.kind == CONTINUOUS_MOUNTAINCAR:
        # continuous_mountain_car.py:157-168: the bonus tests the PRE-transition state
        a = np.clip(action.astype(y.dtype), spec.p("min_action", dt), spec.p("max_action", dt))
        return dt(100.0) * terminal(spec, y).astype(y.dtype) - dt(0.1) * (a * a)
    # acrobot.py:258-270
    return _acrobot_done(y_next).astype(y.dtype) - dt(1.0)


def _acrobot_done(y):
    return (-np.cos(y[0]) - np.cos(y[0] + y[1])) > y.dtype.type(1.0)


def terminal(spec: EnvSpec, y):
    dt = y.dtype.type
    if spec.kind == CARTPOLE:
        # cartpole.py:197-203
        xt = spec.p("x_threshold", dt)
        tt = spec.p("theta_threshold_radians", dt)
        x, theta = y[0], y[2]
        within = (x >= -xt) & (x <= xt) & (theta >= -tt) & (theta <= tt)
        return ~within
    if spec.kind == PENDULUM:
        return np.zeros(y.shape[1], dtype=bool)  # pendulum.py:125-126
    if spec.kind in (MOUNTAINCAR, CONTINUOUS_MOUNTAINCAR):
        # mountain_car.py:182-186, continuous_mountain_car.py:170-174
        return (y[0] >= spec.p("goal_position", dt)) & (y[1] >= spec.p("goal_velocity", dt))
    return _acrobot_done(y)  # acrobot.py:272-277


def truncate(spec: EnvSpec, step_count):
    """TimeLimit.truncate (wrapper/misc.py:193-195); bare envs never truncate
    (base_classic_control.py:74-75)."""
    if spec.max_episode_steps <= 0:
        return np.zeros(step_count.shape, dtype=bool)
    return step_count >= spec.max_episode_steps


def action_bounds(spec: EnvSpec, dtype=np.float32):
    """Box bounds used by PPO.step's clip (ppo.py:228-235); Pendulum: +-max_torque."""
    assert spec.is_box
    if spec.kind == CONTINUOUS_MOUNTAINCAR:
        return spec.p("min_action", dtype), spec.p("max_action", dtype)
    mt = spec.p("max_torque", dtype)
    return -mt, mt
