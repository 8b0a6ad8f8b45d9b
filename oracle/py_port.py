"""Python / NumPy / SciPy restatement of the reference env — TEST INFRASTRUCTURE ONLY.

Second oracle, on the reference's own software stack: it calls the same third-party routines the
reference calls (`scipy.integrate.solve_ivp` with all defaults, `scipy.spatial.transform.Rotation`,
`np.roots`) with operands of the same dtypes, so under one NumPy/SciPy installation it reproduces
the reference bit-for-bit (tests/test_py_port.py pins it to tests/golden/).  It exists because the
reference itself cannot travel to the GPU box: bench.py --impl reference times THIS code (under
oracle/subproc_vec_env.py) as the CPU arm that has the reference's performance character
(Python + SciPy RK45), next to the much faster C oracle.

Restated from /root/reference/my_environment/utils/simulator.py:9-244 (`_Sim`) and
/root/reference/my_environment/envs/rocket_env.py:22-231, 317-402, 503-566, 591-617 (`EnvPort`), plus
the make_env() wrappers of /root/reference/main_6DOF.py:33-53 (`WrappedPort`).  Differences in
form only: one rotation matrix per RHS evaluation instead of three (same values), no history lists,
no rendering.
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import solve_ivp
from scipy.spatial.transform import Rotation

_G0 = 9.81
_RB = 3.66 / 2
_LEN = 40
_ISP = 360
_SREF = np.pi * _RB ** 2
_CA = np.diag([0.82, 0.82, 0.82])
_RHO_EXP = 1 + _G0 * 0.0289644 / 8.3144598 / -0.0065
_R_T = [-15, 0, 0]
_R_CP = [5, 0, 0]


def _cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])


def _height(t, y):
    return y[0]


_height.terminal = True


class _Sim:
    """Rigid-body model + one solve_ivp call per env step (simulator.py:9-104)."""

    def __init__(self, ic, dt):
        self.dt = dt
        self.t = 0
        self.state = ic
        self.state0 = ic
        m = ic[13]                                      # float32 in env mode => float32 inertia under NumPy 2
        self.J = np.diag([.5 * m * _RB ** 2, 1 / 12 * m * (_LEN ** 2 + 3 * _RB ** 2), 1 / 12 * m * (_LEN ** 2 + 3 * _RB ** 2)])
        self.Jinv = np.linalg.inv(self.J)
        self.u = [0, 0, 0]

    @staticmethod
    def thrust_body(u):
        cy, cz, sy, sz = np.cos(u[0]), np.cos(u[1]), np.sin(u[0]), np.sin(u[1])
        m = np.array([[cy * cz, -sy, -cy * sz], [sy * cz, cy, -sy * sz], [sz, 0, cz]])
        return m @ [u[2], 0., 0.]

    @staticmethod
    def rot(q):
        return Rotation.from_quat([q[1], q[2], q[3], q[0]]).as_matrix()

    def rhs(self, t, y, u):
        v, q, w = y[3:6], y[6:10], y[10:13]
        rho = 1.225 * (288.15 / (288.15 + (y[0] - 0) * -0.0065)) ** _RHO_EXP
        R = self.rot(q)
        T = self.thrust_body(u)
        A = -.5 * rho * np.linalg.norm(v) * _SREF * _CA @ (R.transpose() @ v)
        F = R.dot(T + A)
        dv = 1 / y[13] * F + [-_G0, 0, 0]
        wx, wy, wz = w
        om = np.array([[0, -wx, -wy, -wz], [wx, 0, wz, -wy], [wy, -wz, 0, wx], [wz, wy, -wx, 0]])
        dq = 0.5 * om.dot(q)
        tau = _cross(_R_T, T) + _cross(_R_CP, A)
        dw = self.Jinv.dot(tau - np.cross(w, np.dot(self.J, w)))
        dm = -u[2] / (_G0 * _ISP)
        return np.concatenate([v, dv, dq, dw, [dm]])

    def step(self, u):
        sol = solve_ivp(fun=lambda t, y: self.rhs(t, y, u), t_span=[self.t, self.t + self.dt], y0=self.state,
                        events=_height)
        self.state = np.array([c[-1] for c in sol.y])
        self.t = round(self.t + self.dt, 3)
        self.state[6:10] = self.state[6:10] / np.linalg.norm(self.state[6:10])
        self.u = u
        self.nfev = sol.nfev
        return self.state, sol.status


class EnvPort:
    """Rocket6DOF without rendering / plotting; `ep` is rl_rocket_6dof_b200.params.EnvParams."""

    def __init__(self, ep, seed=None):
        self.ep = ep
        self.rng = np.random.RandomState(ep.seed if seed is None else seed)
        self.low, self.high = ep.ic_low, ep.ic_high
        self.max_gimbal = np.deg2rad(20)
        self.max_thrust = 981e3
        self.norm = ep.state_normalizer
        self.b_lo, self.b_hi = ep.bounds_low, ep.bounds_high
        self.c = ep.reward_coeff
        self.sim = None

    def reset(self, ic=None):
        if ic is None:
            ic = self.rng.uniform(low=self.low, high=self.high, size=(14,)).astype(np.float32)
            ic[6:10] = ic[6:10] / np.linalg.norm(ic[6:10])
        self.ic = np.asarray(ic, np.float32)
        self.state = self.ic
        self.sim = _Sim(self.ic, self.ep.timestep)
        return (self.state / self.norm).astype("float32")

    def _oob(self, s):
        r = np.float32(s[0:3])
        return not bool(np.all(r >= self.b_lo) and np.all(r <= self.b_hi))

    def _t_go(self, r, v):
        sol = np.roots([(-9.81) ** 2, 0, -4 * np.linalg.norm(v) ** 2, -24 * np.dot(r, v), -36 * np.linalg.norm(r) ** 2])
        return [z for z in sol if (z.imag == 0 and z.real > 0)][0].real

    def _a_targ(self, r, v, mass):
        g = [-9.81, 0, 0]
        t_go = self._t_go(r, v)
        q = -6 * r / t_go ** 2 - 4 * v / t_go - g
        U = self.max_thrust / mass
        n = np.linalg.norm(q)
        return q if n <= U else q * U / n

    def _v_targ(self, r, v):
        v0 = np.linalg.norm(self.ic[3:6])
        rx = r[0]
        if rx > self.ep.waypoint:
            r_hat, v_hat, tau = r - [self.ep.waypoint, 0, 0], v - [-2, 0, 0], 20
        else:
            r_hat, v_hat, tau = [rx + 1, 0, 0], v - [-1, 0, 0], 100
        t_go = np.linalg.norm(r_hat) / np.linalg.norm(v_hat)
        return -v0 * (np.array(r_hat) / max(1e-3, np.linalg.norm(r_hat))) * (1 - np.exp(-t_go / tau))

    def landing_flags(self, s, eul):
        return {
            "zero_height": s[0] <= 1e-3,
            "velocity_limit": np.linalg.norm(s[3:6]) < self.ep.maximum_v,
            "landing_radius": np.linalg.norm(s[0:3]) < self.ep.target_r,
            "attitude_limit": np.any(abs(eul) < self.ep.land_att_limit),
            "omega_limit": np.any(abs(s[10:13]) < self.ep.omega_lim),
        }

    def step(self, a):
        u = np.float32([a[0] * self.max_gimbal, a[1] * self.max_gimbal, (a[2] + 1) / 2.0 * self.max_thrust])
        self.state, status = self.sim.step(u)
        s = self.state.astype(np.float32)
        eul = Rotation.from_quat(np.roll(s[6:10], -1)).as_euler("zyx")
        oob = self._oob(s)
        done = bool(status) or oob
        r, v, m = s[0:3], s[3:6], s[-1]
        c = self.c
        if self.ep.shaping_type == "acceleration":
            a_t = self._a_targ(np.array(r), np.array(v), m)
            acc = self.sim.rot(self.state[6:10]).dot(self.sim.thrust_body(u)) / m
            shaping = c["alfa"] * np.linalg.norm(acc - a_t)
        else:
            shaping = c["alfa"] * np.linalg.norm(v - self._v_targ(r, v))
        fl = self.landing_flags(s, eul)
        rn, vn = np.linalg.norm(r), np.linalg.norm(v)
        terms = [
            shaping, c["beta"] * u[2], c["eta"],
            c["gamma"] * np.any(np.abs(eul) > self.ep.att_traj_limit),
            c["kappa"] * all(fl.values()),
            max(c["max_r_f"] - rn, 0) * c["w_r_f"],
            max(c["max_v_f"] - vn, 0) * c["w_v_f"] if (rn < c["max_r_f"] and fl["zero_height"]) else 0,
        ]
        reward = sum(terms)
        if oob:
            reward += -50
        self.last = dict(terms=terms, oob=oob, status=status, flags=fl, nfev=self.sim.nfev, u=u)
        return (self.state / self.norm).astype("float32"), reward, done, {"bounds_violation": oob}


class WrappedPort:
    """make_env(): Monitor(TimeLimit(ClipReward(RemoveMassFromObs(env)), 1500)) (main_6DOF.py:44-53)."""

    def __init__(self, ep, seed=None):
        self.env = EnvPort(ep, seed)
        self.max_steps = ep.max_episode_steps
        self.k = 0
        self.ret = 0.0

    def reset(self):
        self.k, self.ret = 0, 0.0
        return self.env.reset()[0:13]

    def step(self, a):
        obs, r, done, info = self.env.step(a)
        r = np.clip(r, -1, 100)
        self.k += 1
        self.ret += r
        if self.k >= self.max_steps and not done:
            info["TimeLimit.truncated"] = True
            done = True
        if done:
            info["episode"] = {"r": float(self.ret), "l": self.k}
        return obs[0:13], r, done, info
