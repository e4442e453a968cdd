"""CartPole-v0 / Acrobot-v1 dynamics, restated from the public gym 0.17.3 classic_control sources
(SURVEY.md Appendix A).  float64 python arithmetic, same operation order as upstream.

Differences from upstream, on purpose:
  * ``math.sin/cos`` (glibc) are used for BOTH envs (upstream Acrobot uses numpy's) so that this
    stand-in, the C oracle (oracle/le_oracle.c) and the CUDA kernels can be compared bit-for-bit up to libm.
  * ``reset_hook``: when set, ``reset()`` takes its uniform draw from ``reset_hook(env)`` instead of
    ``np_random`` so tests can inject the Philox streams the kernels use.
  * no rendering.
"""
import math

import numpy as np

from .. import spaces
from ..core import Env
from ..utils import seeding


class CartPoleEnv(Env):
    def __init__(self):
        self.gravity = 9.8
        self.masscart = 1.0
        self.masspole = 0.1
        self.total_mass = (self.masspole + self.masscart)
        self.length = 0.5  # actually half the pole's length
        self.polemass_length = (self.masspole * self.length)
        self.force_mag = 10.0
        self.tau = 0.02  # seconds between state updates
        self.kinematics_integrator = 'euler'

        self.theta_threshold_radians = 12 * 2 * math.pi / 360
        self.x_threshold = 2.4

        high = np.array([self.x_threshold * 2, np.finfo(np.float32).max,
                         self.theta_threshold_radians * 2, np.finfo(np.float32).max], dtype=np.float32)
        self.action_space = spaces.Discrete(2)
        self.observation_space = spaces.Box(-high, high, dtype=np.float32)

        self.seed()
        self.state = None
        self.steps_beyond_done = None
        self.reset_hook = None

    def seed(self, seed=None):
        self.np_random, seed = seeding.np_random(seed)
        return [seed]

    def step(self, action):
        x, x_dot, theta, theta_dot = self.state
        force = self.force_mag if action == 1 else -self.force_mag
        costheta = math.cos(theta)
        sintheta = math.sin(theta)

        temp = (force + self.polemass_length * theta_dot ** 2 * sintheta) / self.total_mass
        thetaacc = (self.gravity * sintheta - costheta * temp) / (
            self.length * (4.0 / 3.0 - self.masspole * costheta ** 2 / self.total_mass))
        xacc = temp - self.polemass_length * thetaacc * costheta / self.total_mass

        x = x + self.tau * x_dot
        x_dot = x_dot + self.tau * xacc
        theta = theta + self.tau * theta_dot
        theta_dot = theta_dot + self.tau * thetaacc

        self.state = (x, x_dot, theta, theta_dot)

        done = bool(
            x < -self.x_threshold
            or x > self.x_threshold
            or theta < -self.theta_threshold_radians
            or theta > self.theta_threshold_radians
        )

        if not done:
            reward = 1.0
        elif self.steps_beyond_done is None:
            self.steps_beyond_done = 0
            reward = 1.0
        else:
            self.steps_beyond_done += 1
            reward = 0.0

        return np.array(self.state), reward, done, {}

    def reset(self):
        if self.reset_hook is not None:
            self.state = np.asarray(self.reset_hook(self), dtype=np.float64)
        else:
            self.state = self.np_random.uniform(low=-0.05, high=0.05, size=(4,))
        self.steps_beyond_done = None
        return np.array(self.state)


def _wrap(x, m, M):
    diff = M - m
    while x > M:
        x = x - diff
    while x < m:
        x = x + diff
    return x


def _bound(x, m, M):
    return min(max(x, m), M)


class AcrobotEnv(Env):
    dt = .2

    LINK_LENGTH_1 = 1.
    LINK_LENGTH_2 = 1.
    LINK_MASS_1 = 1.
    LINK_MASS_2 = 1.
    LINK_COM_POS_1 = 0.5
    LINK_COM_POS_2 = 0.5
    LINK_MOI = 1.

    MAX_VEL_1 = 4 * math.pi
    MAX_VEL_2 = 9 * math.pi

    AVAIL_TORQUE = [-1., 0., +1]
    torque_noise_max = 0.
    book_or_nips = "book"

    def __init__(self):
        high = np.array([1.0, 1.0, 1.0, 1.0, self.MAX_VEL_1, self.MAX_VEL_2], dtype=np.float32)
        self.observation_space = spaces.Box(low=-high, high=high, dtype=np.float32)
        self.action_space = spaces.Discrete(3)
        self.state = None
        self.reset_hook = None
        self.seed()

    def seed(self, seed=None):
        self.np_random, seed = seeding.np_random(seed)
        return [seed]

    def reset(self):
        if self.reset_hook is not None:
            self.state = np.asarray(self.reset_hook(self), dtype=np.float64)
        else:
            self.state = self.np_random.uniform(low=-0.1, high=0.1, size=(4,))
        return self._get_ob()

    def step(self, a):
        s = [float(v) for v in self.state]
        torque = self.AVAIL_TORQUE[a]
        y0 = s + [torque]
        ns = self._rk4_step(y0, self.dt)
        ns = ns[:4]
        ns[0] = _wrap(ns[0], -math.pi, math.pi)
        ns[1] = _wrap(ns[1], -math.pi, math.pi)
        ns[2] = _bound(ns[2], -self.MAX_VEL_1, self.MAX_VEL_1)
        ns[3] = _bound(ns[3], -self.MAX_VEL_2, self.MAX_VEL_2)
        self.state = np.array(ns)
        terminal = self._terminal()
        reward = -1. if not terminal else 0.
        return (self._get_ob(), reward, terminal, {})

    def _get_ob(self):
        s = self.state
        return np.array([math.cos(s[0]), math.sin(s[0]), math.cos(s[1]), math.sin(s[1]), s[2], s[3]])

    def _terminal(self):
        s = self.state
        return bool(-math.cos(s[0]) - math.cos(s[1] + s[0]) > 1.)

    def _dsdt(self, s_augmented):
        m1 = self.LINK_MASS_1
        m2 = self.LINK_MASS_2
        l1 = self.LINK_LENGTH_1
        lc1 = self.LINK_COM_POS_1
        lc2 = self.LINK_COM_POS_2
        I1 = self.LINK_MOI
        I2 = self.LINK_MOI
        g = 9.8
        pi = math.pi
        cos = math.cos
        sin = math.sin
        a = s_augmented[-1]
        s = s_augmented[:-1]
        theta1 = s[0]
        theta2 = s[1]
        dtheta1 = s[2]
        dtheta2 = s[3]
        d1 = m1 * lc1 ** 2 + m2 * (l1 ** 2 + lc2 ** 2 + 2 * l1 * lc2 * cos(theta2)) + I1 + I2
        d2 = m2 * (lc2 ** 2 + l1 * lc2 * cos(theta2)) + I2
        phi2 = m2 * lc2 * g * cos(theta1 + theta2 - pi / 2.)
        phi1 = - m2 * l1 * lc2 * dtheta2 ** 2 * sin(theta2) \
            - 2 * m2 * l1 * lc2 * dtheta2 * dtheta1 * sin(theta2) \
            + (m1 * lc1 + m2 * l1) * g * cos(theta1 - pi / 2) + phi2
        # "book" variant (gym 0.17.3 default)
        ddtheta2 = (a + d2 / d1 * phi1 - m2 * l1 * lc2 * dtheta1 ** 2 * sin(theta2) - phi2) \
            / (m2 * lc2 ** 2 + I2 - d2 ** 2 / d1)
        ddtheta1 = -(d2 * ddtheta2 + phi1) / d1
        return [dtheta1, dtheta2, ddtheta1, ddtheta2, 0.]

    def _rk4_step(self, y0, dt):
        # one classical RK4 step as gym's rk4(derivs, y0, [0, dt]) does elementwise on float64 arrays
        dt2 = dt / 2.0
        n = len(y0)
        k1 = self._dsdt(y0)
        k2 = self._dsdt([y0[i] + dt2 * k1[i] for i in range(n)])
        k3 = self._dsdt([y0[i] + dt2 * k2[i] for i in range(n)])
        k4 = self._dsdt([y0[i] + dt * k3[i] for i in range(n)])
        return [y0[i] + dt / 6.0 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]) for i in range(n)]
