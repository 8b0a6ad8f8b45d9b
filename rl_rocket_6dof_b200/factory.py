"""Factories mirroring main_6DOF.py:44-53 (`make_env`) for the batched env."""
from __future__ import annotations

from typing import Optional

from .params import load_config


def make_vec_env(num_envs: int, config_path: Optional[str] = None, **kw):
    """Monitor(TimeLimit(ClipReward(RemoveMassFromObs(Rocket6DOF)))) x num_envs, as one VecEnv."""
    from .vec_env import Rocket6DOFVecEnv
    sb3_config, env_config = load_config(config_path)
    return Rocket6DOFVecEnv(num_envs, env_config, sb3_config, **kw)


def make_env(config_path: Optional[str] = None, **kw):
    """The reference's single wrapped env == a 1-env VecEnv (SB3 wraps single envs the same way)."""
    return make_vec_env(1, config_path, **kw)


def make_annealed_vec_env(num_envs: int, config_path: Optional[str] = None, **kw):
    """`make_annealed_env()` of main_6DOF.py:55-69, batched: Monitor(TimeLimit(RewardAnnealing(RemoveMassFromObs(env))))
    — no ClipReward, reward rebuilt by RewardAnnealing (wrappers.py:39-61)."""
    return make_vec_env(num_envs, config_path, clip_reward=False, reward_annealing=True, **kw)


def make_annealed_env(config_path: Optional[str] = None, **kw):
    return make_annealed_vec_env(1, config_path, **kw)
