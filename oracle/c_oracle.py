"""ctypes front-end of oracle/r6_oracle.c — TEST INFRASTRUCTURE ONLY (see the C file's header).

Only tests/, `__graft_entry__.smoke()` and bench.py's CPU-baseline legs import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libr6oracle.so")


class R6OParams(C.Structure):
    _fields_ = [
        ("dt", C.c_double), ("max_gimbal", C.c_double),
        ("max_thrust", C.c_float), ("beta", C.c_float), ("w_v_f", C.c_float), ("w_r_f", C.c_float),
        ("max_r_f", C.c_float), ("max_v_f", C.c_float), ("maximum_v", C.c_float),
        ("target_r", C.c_float), ("zero_height_tol", C.c_float),
        ("bounds_low", C.c_float * 3), ("bounds_high", C.c_float * 3),
        ("normalizer", C.c_double * 14),
        ("alfa", C.c_double), ("eta", C.c_double), ("gamma", C.c_double), ("kappa", C.c_double),
        ("att_traj_limit", C.c_double * 3), ("land_att_limit", C.c_double * 3),
        ("omega_lim", C.c_double * 3), ("waypoint", C.c_double),
        ("shaping_velocity", C.c_int32), ("n_t", C.c_int32),
        ("t_table", C.POINTER(C.c_double)),
    ]


class R6OEnv(C.Structure):
    _fields_ = [("y", C.c_double * 14), ("m0", C.c_float), ("v0", C.c_float),
                ("k", C.c_int32), ("pad", C.c_int32)]


class R6OOut(C.Structure):
    _fields_ = [
        ("state", C.c_double * 14), ("obs", C.c_float * 14), ("reward", C.c_double),
        ("terms", C.c_double * 7), ("u", C.c_float * 3),
        ("done", C.c_int32), ("oob", C.c_int32), ("status", C.c_int32), ("nfev", C.c_int32),
        ("flags", C.c_int32 * 5), ("tgo_npos", C.c_int32),
    ]


ENV_DTYPE = np.dtype(R6OEnv)
OUT_DTYPE = np.dtype(R6OOut)


def build(force: bool = False) -> str:
    """Compiles r6_oracle.c with gcc (oracle/Makefile). Building the checker is not using it."""
    src = os.path.join(HERE, "r6_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"] + (["-B"] if force else []))
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        assert L.r6o_sizeof_params() == C.sizeof(R6OParams), "R6OParams layout mismatch"
        assert L.r6o_sizeof_env() == C.sizeof(R6OEnv) == ENV_DTYPE.itemsize, "R6OEnv layout mismatch"
        assert L.r6o_sizeof_out() == C.sizeof(R6OOut) == OUT_DTYPE.itemsize, "R6OOut layout mismatch"
        L.r6o_tgo.restype = C.c_double
        L.r6o_tgo.argtypes = [C.c_double] * 4 + [C.POINTER(C.c_int)]
        L.r6o_sim_step_raw.restype = C.c_int
        L.r6o_sim_step_raw.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double,
                                       C.c_double, C.c_double, C.POINTER(C.c_int)]
        L.r6o_env_step_batch.argtypes = [C.POINTER(R6OParams), C.c_void_p, C.c_int64, C.c_void_p,
                                         C.c_void_p, C.c_int]
        L.r6o_reset_from_sample.argtypes = [C.POINTER(R6OParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.r6o_euler_zyx.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.r6o_denormalize_action.argtypes = [C.POINTER(R6OParams), C.c_void_p, C.c_void_p]
        L.r6o_step_consts.argtypes = [C.c_float, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def make_params(ep) -> R6OParams:
    """EnvParams (rl_rocket_6dof_b200.params) -> the oracle's parameter struct."""
    rc = ep.reward_coeff
    p = R6OParams()
    p.dt = float(ep.timestep)
    p.max_gimbal = float(ep.max_gimbal)
    p.max_thrust = np.float32(ep.max_thrust)
    p.beta = np.float32(rc["beta"])
    p.w_v_f, p.w_r_f = np.float32(rc["w_v_f"]), np.float32(rc["w_r_f"])
    p.max_r_f, p.max_v_f = np.float32(rc["max_r_f"]), np.float32(rc["max_v_f"])
    p.maximum_v, p.target_r = np.float32(ep.maximum_v), np.float32(ep.target_r)
    p.zero_height_tol = np.float32(1e-3)
    p.bounds_low[:] = [float(x) for x in ep.bounds_low]
    p.bounds_high[:] = [float(x) for x in ep.bounds_high]
    p.normalizer[:] = [float(x) for x in ep.state_normalizer]
    p.alfa, p.eta, p.gamma, p.kappa = float(rc["alfa"]), float(rc["eta"]), float(rc["gamma"]), float(rc["kappa"])
    p.att_traj_limit[:] = [float(x) for x in ep.att_traj_limit]
    p.land_att_limit[:] = [float(x) for x in ep.land_att_limit]
    p.omega_lim[:] = [float(x) for x in ep.omega_lim]
    p.waypoint = float(ep.waypoint)
    p.shaping_velocity = 1 if ep.shaping_type == "velocity" else 0
    tt = np.ascontiguousarray(ep.t_table, np.float64)
    p._keep = tt
    p.n_t = len(tt)
    p.t_table = tt.ctypes.data_as(C.POINTER(C.c_double))
    return p


class OracleBatch:
    """N independent oracle envs with the Rocket6DOF reset/step contract (no RNG: initial
    conditions are injected, exactly as the parity plan of SURVEY.md §8c prescribes)."""

    def __init__(self, ep, n: int, nthreads: int = 1):
        self.ep = ep
        self.p = make_params(ep)
        self.n = n
        self.nthreads = nthreads
        self.envs = np.zeros(n, ENV_DTYPE)
        self.outs = np.zeros(n, OUT_DTYPE)

    def reset_from_samples(self, samples: np.ndarray, idx=None):
        """samples: float32 [m,14] raw Box samples (quaternion not yet normalised)."""
        L = lib()
        samples = np.ascontiguousarray(samples, np.float32).reshape(-1, 14)
        idx = np.arange(self.n) if idx is None else np.atleast_1d(idx)
        obs = np.zeros((len(idx), 14), np.float32)
        ic = np.zeros((len(idx), 14), np.float32)
        for j, i in enumerate(idx):
            L.r6o_reset_from_sample(C.byref(self.p), samples[j].ctypes.data,
                                    self.envs[i:i + 1].ctypes.data, obs[j].ctypes.data, ic[j].ctypes.data)
        return obs, ic

    def set_state(self, y: np.ndarray, m0, k, v0=None, idx=None):
        idx = np.arange(self.n) if idx is None else np.atleast_1d(idx)
        self.envs["y"][idx] = y
        self.envs["m0"][idx] = m0
        self.envs["k"][idx] = k
        if v0 is not None:
            self.envs["v0"][idx] = v0

    def step(self, actions: np.ndarray):
        a = np.ascontiguousarray(actions, np.float32).reshape(self.n, 3)
        lib().r6o_env_step_batch(C.byref(self.p), self.envs.ctypes.data, self.n, a.ctypes.data,
                                 self.outs.ctypes.data, self.nthreads)
        return self.outs


def tgo(c2, c3, c4, c0=(-9.81) ** 2):
    npos = C.c_int(0)
    t = lib().r6o_tgo(c0, float(c2), float(c3), float(c4), C.byref(npos))
    return t, npos.value


def sim_step_raw(y, u, m0, t, dt):
    yy = np.array(y, np.float64)
    uu = np.array(u, np.float64)
    nfev = C.c_int(0)
    st = lib().r6o_sim_step_raw(yy.ctypes.data_as(C.POINTER(C.c_double)), uu.ctypes.data_as(C.POINTER(C.c_double)),
                                float(m0), float(t), float(dt), C.byref(nfev))
    return yy, st, nfev.value


def euler_zyx(q):
    qq = np.array(q, np.float64)
    e = np.zeros(3)
    lib().r6o_euler_zyx(qq.ctypes.data_as(C.POINTER(C.c_double)), e.ctypes.data_as(C.POINTER(C.c_double)))
    return e


def denormalize_action(p: R6OParams, a):
    aa = np.array(a, np.float32)
    u = np.zeros(3, np.float32)
    lib().r6o_denormalize_action(C.byref(p), aa.ctypes.data, u.ctypes.data)
    return u


def step_consts(m0, u):
    uu = np.array(u, np.float32)
    out = np.zeros(8)
    lib().r6o_step_consts(np.float32(m0), uu.ctypes.data, out.ctypes.data)
    return out
