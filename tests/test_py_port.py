"""The Python/SciPy port (oracle/py_port.py, the CPU arm of bench.py --impl reference) against the
fixtures dumped from the reference.  Same third-party calls on the same dtypes => bit-exact here."""
import warnings

import numpy as np

from oracle.py_port import EnvPort
from parity_utils import env_params, golden

warnings.filterwarnings("ignore")


def _replay(ep, rec, T):
    env = EnvPort(ep)
    ic_at = {int(s): rec["ic"][j] for j, s in enumerate(rec["ic_step"]) if s >= 0}
    for k in range(T):
        if k in ic_at:
            env.reset(ic_at[k])
        obs, r, done, info = env.step(rec["action"][k])
        assert np.array_equal(env.state, rec["state"][k]), k
        assert np.array_equal(obs, rec["obs"][k]), k
        assert r == rec["reward"][k] and done == rec["done"][k], k
        assert env.last["nfev"] == rec["nfev"][k] and env.last["status"] == rec["status"][k]
        assert [bool(v) for v in env.last["flags"].values()] == list(rec["flags"][k])
        assert np.array_equal(np.array([float(x) for x in env.last["terms"]]), rec["terms"][k])


def test_py_port_bit_exact_config1_prefix():
    _replay(env_params(), golden("config1"), 300)


def test_py_port_bit_exact_velocity_prefix():
    _replay(env_params(reward_shaping_type="velocity"), golden("velocity"), 200)


def test_subproc_vec_env_protocol():
    from oracle.subproc_vec_env import SubprocVecEnvPort
    vec = SubprocVecEnvPort(2)
    try:
        obs = vec.reset()
        assert obs.shape == (2, 13) and obs.dtype == np.float32
        for _ in range(5):
            obs, rew, done, info = vec.step(np.zeros((2, 3), np.float32))
        assert obs.shape == (2, 13) and rew.shape == (2,) and done.shape == (2,) and len(info) == 2
        assert np.all(rew <= 100) and np.all(rew >= -1)
    finally:
        vec.close()
