"""CPU arm: the UNMODIFIED reference env (baseline/_ref, installed by baseline/install_ref.py) built exactly as
`make_env()` of /root/reference/main_6DOF.py:44-53 —

    Monitor(TimeLimit(ClipReward(RemoveMassFromObs(gym.make("my_environment/Falcon6DOF-v0", **env_config))), 1500))

— and stepped under a SubprocVecEnv-protocol harness (oracle/subproc_vec_env.py), one worker process per host core,
as BASELINE.md §3 prescribes.  BASELINE / TEST INFRASTRUCTURE ONLY: nothing under rl_rocket_6dof_b200/ imports this.

What is NOT the reference's own code here, and why:
  * `gym`, `gym.wrappers.TimeLimit`, `stable_baselines3.common.monitor.Monitor`, `pyvista`, `matplotlib`, `plotly`
    are import shims (oracle/ref_shims): the real packages are not installed and there is no network;
  * `ClipReward` is restated from main_6DOF.py:33-42 (7 lines): main_6DOF.py itself cannot be imported (it pulls
    stable_baselines3.PPO and wandb.integration.sb3 at module level and reads config.yaml from the cwd);
  * SubprocVecEnv is the pipe-per-worker stand-in with SB3's worker protocol.
Everything that does arithmetic — Rocket6DOF, Simulator6DOF, SciPy solve_ivp, RemoveMassFromObs — is stock.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
SHIMS = os.path.join(ROOT, "oracle", "ref_shims")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "my_environment", "envs", "rocket_env.py"))


def _import_reference():
    if not available():
        raise RuntimeError("baseline/_ref is missing: run `python baseline/install_ref.py` where /root/reference exists")
    for p in (SHIMS, REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    warnings.filterwarnings("ignore")
    import gym                                      # shim
    import my_environment                           # noqa: F401  (registers my_environment/Falcon6DOF-v0)
    from my_environment.wrappers import RemoveMassFromObs
    from gym.wrappers import TimeLimit              # shim
    from stable_baselines3.common.monitor import Monitor   # shim
    return gym, RemoveMassFromObs, TimeLimit, Monitor


def load_config():
    """main_6DOF.py:18-27 on baseline/_ref/config.yaml."""
    import yaml
    with open(os.path.join(REF, "config.yaml")) as f:
        config = yaml.safe_load(f)
    return config["sb3_config"], config["env_config"]


def make_env(seed=None):
    """main_6DOF.py:44-53.  `seed` (SubprocVecEnv.seed gives worker i seed + i) re-seeds the IC sampler through the
    env's own seed() (rocket_env.py:619-621)."""
    gym, RemoveMassFromObs, TimeLimit, Monitor = _import_reference()
    sb3_config, env_config = load_config()
    max_episode_steps = int(sb3_config["max_time"] / env_config["timestep"])          # main_6DOF.py:31

    class ClipReward(gym.RewardWrapper):                                             # main_6DOF.py:33-42
        def __init__(self, env, min_reward=-1, max_reward=100):
            super().__init__(env)
            self.min_reward = min_reward
            self.max_reward = max_reward
            self.reward_range = (min_reward, max_reward)

        def reward(self, reward):
            return np.clip(reward, self.min_reward, self.max_reward)

    kwargs = env_config
    env = ClipReward(RemoveMassFromObs(gym.make("my_environment/Falcon6DOF-v0", **kwargs)))
    env = TimeLimit(env, max_episode_steps=max_episode_steps)
    env = Monitor(env)
    if seed is not None:
        env.seed(seed)
    return env


def time_reference(n_workers: int, steps: int, warmup: int, seed0: int = 42) -> dict:
    """env-steps/s of the reference env under the SubprocVecEnv harness: `steps` VecEnv.step calls over `n_workers`
    single-env worker processes after `warmup` untimed ones.  Actions: np.random.default_rng(worker).uniform(-1,1,3)
    float32; auto-reset on."""
    import time
    sys.path.insert(0, ROOT)
    from oracle.subproc_vec_env import SubprocVecEnvPort
    vec = SubprocVecEnvPort(n_workers, seed0=seed0, env_fn=make_env)
    try:
        vec.reset()
        rngs = [np.random.default_rng(i) for i in range(n_workers)]

        def acts():
            return [r.uniform(-1, 1, 3).astype(np.float32) for r in rngs]
        for _ in range(warmup):
            vec.step(acts())
        marks = [time.perf_counter()]
        episodes = 0
        for k in range(steps):
            _, _, dones, _ = vec.step(acts())
            episodes += int(np.sum(dones))
            marks.append(time.perf_counter())
        dt = marks[-1] - marks[0]
    finally:
        vec.close()
    return {"steps_per_s": n_workers * steps / dt, "seconds": dt, "vec_steps": steps, "episodes": episodes,
            "marks": marks,
            "sample": f"baseline/_ref (unmodified reference Rocket6DOF, make_env() wrappers of main_6DOF.py:44-53), "
                      f"{n_workers} SubprocVecEnv-protocol worker processes x {steps} steps after {warmup} warm-up, "
                      f"random actions, auto-reset"}


if __name__ == "__main__":
    import json
    n = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
    out = time_reference(n, int(sys.argv[2]) if len(sys.argv) > 2 else 200, 20)
    out.pop("marks")
    print(json.dumps(out))
