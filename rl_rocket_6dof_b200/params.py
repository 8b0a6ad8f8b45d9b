"""Host-side derivation of the env constants (config.yaml -> R6Params).

Restates `Rocket6DOF.__init__` (/root/reference/my_environment/envs/rocket_env.py:71-134,159-168)
and the `make_env()` wrapper parameters (/root/reference/main_6DOF.py:29-53).  The constants are
computed with NumPy expressions of the same operand types as the reference's, because the values
depend on NumPy's promotion rules (`t_free_fall`, `v_max`, `omega_max` carry
float32 rounding under NumPy >= 2); the kernel only ever sees the resulting numbers.
`tests/test_abi.py::test_params_match_reference_constants` pins them to `tests/golden/constants.npz` (dumped from the reference).
"""
from __future__ import annotations

import copy
import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

NSTATE = 14
STATE_NAMES = ["x", "y", "z", "vx", "vy", "vz", "q0", "q1", "q2", "q3", "omega1", "omega2", "omega3", "mass"]
ACTION_NAMES = ["gimbal_y", "gimbal_z", "thrust"]
REWARD_TERMS_ACC = ["atarg_tracking", "thrust_penalty", "eta", "attitude_constraint",
                    "goal_conditions", "final_position", "final_velocity"]
REWARD_TERMS_VEL = ["vtarg_tracking"] + REWARD_TERMS_ACC[1:]

_DEFAULT_CONFIG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config.yaml")


def load_config(path: str | None = None):
    """Same contract as `load_config()` of main_6DOF.py:18-27: returns (sb3_config, env_config)."""
    import yaml

    with open(path or _DEFAULT_CONFIG) as f:
        cfg = yaml.safe_load(f)
    return cfg["sb3_config"], cfg["env_config"]


class R6Params(C.Structure):
    """ctypes mirror of `struct R6Params` in include/r6dof.h (layout checked at load time)."""

    _fields_ = [
        ("dt", C.c_double),
        ("max_gimbal", C.c_double),
        ("normalizer", C.c_double * NSTATE),
        ("alfa", C.c_double), ("eta", C.c_double), ("gamma", C.c_double), ("kappa", C.c_double),
        ("att_traj_limit", C.c_double * 3),
        ("land_att_limit", C.c_double * 3),
        ("omega_lim", C.c_double * 3),
        ("waypoint", C.c_double),
        ("clip_lo", C.c_double), ("clip_hi", C.c_double),
        ("oob_penalty", C.c_double),
        ("max_thrust", C.c_float),
        ("beta", C.c_float), ("w_v_f", C.c_float), ("w_r_f", C.c_float),
        ("max_r_f", C.c_float), ("max_v_f", C.c_float),
        ("maximum_v", C.c_float), ("target_r", C.c_float), ("zero_height_tol", C.c_float),
        ("bounds_low", C.c_float * 3), ("bounds_high", C.c_float * 3),
        ("ic_low", C.c_float * NSTATE), ("ic_high", C.c_float * NSTATE),
        ("shaping_velocity", C.c_int32),
        ("max_episode_steps", C.c_int32),
        ("clip_reward", C.c_int32),
        ("auto_reset", C.c_int32),
        ("n_t", C.c_int32),
        ("obs_rows", C.c_int32),
        ("precision", C.c_int32),
        ("reward_mode", C.c_int32),
        ("va_threshold", C.c_double),
        ("va_weight", C.c_double),
        ("xi", C.c_float),
        ("obs_row_major", C.c_int32),
    ]


@dataclass
class EnvParams:
    """All numbers the step needs, as NumPy scalars/arrays of the reference's own dtypes."""

    timestep: float
    seed: int
    ic_low: np.ndarray          # float32[14]  init_space.low   (rocket_env.py:78-81)
    ic_high: np.ndarray         # float32[14]
    max_gimbal: float           # np.deg2rad(20)                 (:86)
    max_thrust: float           # 981e3                          (:87)
    state_normalizer: np.ndarray  # float64[14]                  (:106-126)
    bounds_low: np.ndarray      # float32[3]                     (:129-134)
    bounds_high: np.ndarray
    att_traj_limit: np.ndarray  # float64[3] rad                 (:159)
    land_att_limit: np.ndarray  # float64[3] rad                 (:165)
    omega_lim: np.ndarray       # float64[3] hard-coded 0.2      (:167)
    waypoint: float             # (:168) required key
    target_r: float             # (:162)
    maximum_v: float            # (:163)
    reward_coeff: dict
    shaping_type: str
    max_episode_steps: int      # TimeLimit, main_6DOF.py:31
    clip_reward: bool = True    # ClipReward(-1, 100), main_6DOF.py:33-42
    clip_lo: float = -1.0
    clip_hi: float = 100.0
    t_table: np.ndarray = field(default=None, repr=False)   # float64[max_episode_steps + 2]

    @property
    def reward_term_names(self):
        return REWARD_TERMS_VEL if self.shaping_type == "velocity" else REWARD_TERMS_ACC

    def to_struct(self, auto_reset: bool = True, clip_reward: bool | None = None,
                  time_limit: bool = True, obs_rows: int = 0, precision: int = 0, reward_annealing: bool = False,
                  vertical_attitude_reward=None, obs_row_major: bool = False) -> R6Params:
        """reward_annealing: RewardAnnealing wrapper (xi = reward_coeff.get("xi", 0.01), wrappers.py:42);
        vertical_attitude_reward: None or (threshold_height, weight) of VerticalAttitudeReward (wrappers.py:129)."""
        p = R6Params()
        rc = self.reward_coeff
        p.dt = float(self.timestep)
        p.max_gimbal = float(self.max_gimbal)
        p.normalizer[:] = [float(x) for x in self.state_normalizer]
        p.alfa, p.eta, p.gamma, p.kappa = float(rc["alfa"]), float(rc["eta"]), float(rc["gamma"]), float(rc["kappa"])
        p.att_traj_limit[:] = [float(x) for x in self.att_traj_limit]
        p.land_att_limit[:] = [float(x) for x in self.land_att_limit]
        p.omega_lim[:] = [float(x) for x in self.omega_lim]
        p.waypoint = float(self.waypoint)
        p.clip_lo, p.clip_hi = float(self.clip_lo), float(self.clip_hi)
        p.oob_penalty = -50.0
        # python scalars that meet np.float32 operands are "weak" => used as float32
        p.max_thrust = np.float32(self.max_thrust)
        p.beta = np.float32(rc["beta"])
        p.w_v_f, p.w_r_f = np.float32(rc["w_v_f"]), np.float32(rc["w_r_f"])
        p.max_r_f, p.max_v_f = np.float32(rc["max_r_f"]), np.float32(rc["max_v_f"])
        p.maximum_v, p.target_r = np.float32(self.maximum_v), np.float32(self.target_r)
        p.zero_height_tol = np.float32(1e-3)
        p.bounds_low[:] = [np.float32(x) for x in self.bounds_low]
        p.bounds_high[:] = [np.float32(x) for x in self.bounds_high]
        p.ic_low[:] = [np.float32(x) for x in self.ic_low]
        p.ic_high[:] = [np.float32(x) for x in self.ic_high]
        p.shaping_velocity = 1 if self.shaping_type == "velocity" else 0
        p.max_episode_steps = int(self.max_episode_steps) if time_limit else 0
        p.clip_reward = int(self.clip_reward if clip_reward is None else clip_reward)
        p.auto_reset = int(auto_reset)
        p.n_t = int(len(self.t_table))
        p.obs_rows = int(obs_rows)
        p.precision = int(precision)
        p.reward_mode = (1 if reward_annealing else 0) | (2 if vertical_attitude_reward is not None else 0)
        p.xi = np.float32(rc.get("xi", 0.01))
        th, wt = vertical_attitude_reward if vertical_attitude_reward is not None else (1e-3, -0.5)
        p.va_threshold, p.va_weight = float(th), float(wt)
        p.obs_row_major = int(bool(obs_row_major))
        return p


def make_t_table(timestep: float, n: int) -> np.ndarray:
    """t_0 = 0, t_k = round(t_{k-1} + dt, 3) — the clock of Simulator6DOF (simulator.py:13,92)."""
    t = np.zeros(n, np.float64)
    cur = 0
    for k in range(1, n):
        cur = round(cur + timestep, 3)
        t[k] = cur
    return t


def derive_params(env_config: dict, sb3_config: dict | None = None, *, clip_reward: bool = True,
                  max_episode_steps: int | None = None) -> EnvParams:
    """config.yaml's `env_config` (+ `sb3_config.max_time`) -> EnvParams."""
    cfg = copy.deepcopy(env_config)
    ic_mean = np.float32(cfg.get("IC", [500, 100, 100, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 50e3]))
    ic_range = np.float32(cfg.get("ICRange", [50, 10, 10, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 1e3]))
    timestep = cfg.get("timestep", 0.1)
    # gym's Box stores low/high in float32
    ic_low = (ic_mean - ic_range / 2).astype(np.float32)
    ic_high = (ic_mean + ic_range / 2).astype(np.float32)

    max_gimbal = np.deg2rad(20)
    max_thrust = 981e3

    # np.float32 element (op) python float stays float32 under NEP-50, as in the reference
    t_free_fall = (-ic_mean[3] + np.sqrt(ic_mean[3] ** 2 + 2 * 9.81 * ic_mean[0])) / 9.81
    inertia = 6.04e6
    lever_arm = 15.0
    omega_max = max_thrust * np.sin(max_gimbal) * lever_arm / inertia * t_free_fall / 5.0
    v_max = 2 * 9.81 * t_free_fall
    normalizer = np.maximum(
        np.array([
            1.2 * abs(ic_mean[0]), 1.5 * abs(ic_mean[1]), 1.5 * abs(ic_mean[2]),
            v_max, v_max, v_max,
            1.1, 1.1, 1.1, 1.1,
            omega_max, omega_max, omega_max,
            ic_mean[13] + ic_range[13],
        ]),
        1,
    )
    hi = 0.9 * np.maximum(normalizer[0:3], 200)
    lo = np.insert(-0.9 * np.maximum(normalizer[1:3], 200), 0, -30)

    traj = cfg.get("trajectory_limits", {"attitude_limit": [85, 85, 360]})
    land = cfg.get("landing_params", None)
    if land is None or "waypoint" not in land:
        # the reference raises KeyError here too (rocket_env.py:168)
        raise KeyError("waypoint")
    coeff = cfg.get("reward_coeff", {
        "alfa": -0.01, "beta": -1e-8, "eta": 2, "gamma": -10, "delta": -5, "kappa": 10,
        "w_r_f": 1, "w_v_f": 5, "max_r_f": 100, "max_v_f": 100})
    if max_episode_steps is None:
        max_time = (sb3_config or {}).get("max_time", 150)
        max_episode_steps = int(max_time / timestep)

    return EnvParams(
        timestep=timestep,
        seed=int(cfg.get("seed", 42)),
        ic_low=ic_low, ic_high=ic_high,
        max_gimbal=float(max_gimbal), max_thrust=max_thrust,
        state_normalizer=np.asarray(normalizer, np.float64),
        bounds_low=lo.astype(np.float32), bounds_high=hi.astype(np.float32),
        att_traj_limit=np.deg2rad(traj["attitude_limit"]).astype(np.float64),
        land_att_limit=np.deg2rad(land["landing_attitude_limit"]).astype(np.float64),
        omega_lim=np.array([0.2, 0.2, 0.2]),
        waypoint=float(land["waypoint"]),
        target_r=land["landing_radius"], maximum_v=land["maximum_velocity"],
        reward_coeff=dict(coeff),
        shaping_type=cfg.get("reward_shaping_type", "acceleration"),
        max_episode_steps=int(max_episode_steps),
        clip_reward=clip_reward,
        t_table=make_t_table(timestep, int(max_episode_steps) + 2),
    )
