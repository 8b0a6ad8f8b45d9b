"""Shared helpers of the parity tests: golden loading, tolerances, replay."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

RTOL_STATE = 1e-9          # fp64 path, north_star
FLOOR_FRAC = 1e-3          # s_i = normalizer_i * 1e-3
RTOL_REWARD_UNIT = 1e-9    # reward evaluated on the same state
RTOL_REWARD_TRAJ = 1e-6    # reward along a trajectory (float32 casts inside the reward)
REWARD_FLOOR = 1e-3
# Along a trajectory r, v, m, q are cast to float32 before the reward (rocket_env.py:206): a 1e-13
# state difference occasionally flips a float32 ulp and moves individual reward TERMS (|term| ~ 0.1)
# by ~5e-9 absolute.  The sum of terms can cancel to ~1e-3, so the trajectory-level error is
# measured against max(|reward|, 0.1) (0.1 = the constant `eta` term of config.yaml).
REWARD_FLOOR_TRAJ = 0.1


_CACHE = {}


def golden(name):
    """Loads a fixture once, fully decompressed (NpzFile re-reads the archive on every access)."""
    if name not in _CACHE:
        with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
            _CACHE[name] = {k: z[k] for k in z.files}
    return _CACHE[name]


def env_params(**over):
    from rl_rocket_6dof_b200.params import derive_params, load_config
    sb3, env = load_config()
    env = {**env, **over}
    return derive_params(env, sb3)


def state_err(x, ref, normalizer):
    """max over components of |x-ref| / max(|ref|, s_i) — compare with RTOL_STATE."""
    s = np.asarray(normalizer) * FLOOR_FRAC
    return np.max(np.abs(x - ref) / np.maximum(np.abs(ref), s), axis=-1)


def reward_err(x, ref, floor=REWARD_FLOOR):
    return np.abs(x - ref) / np.maximum(np.abs(ref), floor)


def reward_err_traj(x, ref):
    return reward_err(x, ref, REWARD_FLOOR_TRAJ)


def f32_ulp_diff(a, b):
    """|a-b| in units of float32 ulps of b (element-wise)."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    ulp = np.spacing(np.maximum(np.abs(b), np.float32(1e-3)))   # floor: obs are O(1) normalised
    return np.abs(a.astype(np.float64) - b.astype(np.float64)) / ulp


def f32_norm3(v):
    """OpenBLAS sdot rule: f32 products, f64 accumulation, one f32 rounding."""
    v = np.asarray(v, np.float32)
    acc = np.float64(0)
    for x in v:
        acc += np.float64(np.float32(x * x))
    return np.sqrt(np.float32(acc))
