"""Rocket6DOF — single-environment facade with the reference's gym contract
(/root/reference/my_environment/envs/rocket_env.py:16-231) on top of a 1-env CUDA batch.

Same constructor keywords as the reference (`config.yaml`'s env_config is splatted into it), same
`reset() -> obs(14,) float32`, `step(a) -> (obs, reward, done, info)`, `seed`, `get_state`,
`used_mass`, `state_normalizer`, `observation_space` / `action_space`; the `info` dict carries
`rewards_dict`, `is_done`, `state_history`, `action_history`, `timesteps`, `bounds_violation` and the
individual reward terms like the reference (:217-226).  Rendering / plotting are stripped.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .batch import F_EVENT, F_OOB, Rocket6DOFBatch
from .params import ACTION_NAMES, STATE_NAMES, derive_params
from .spaces import make_box


class Rocket6DOF:
    metadata = {"render.modes": [], "render_fps": 10}

    def __init__(self, IC=None, ICRange=None, timestep=0.1, seed=42, reward_shaping_type="acceleration",
                 reward_coeff=None, trajectory_limits=None, landing_params=None, *, device="cuda"):
        cfg = dict(timestep=timestep, seed=seed, reward_shaping_type=reward_shaping_type)
        for k, v in (("IC", IC), ("ICRange", ICRange), ("reward_coeff", reward_coeff),
                     ("trajectory_limits", trajectory_limits), ("landing_params", landing_params)):
            if v is not None:
                cfg[k] = v
        if landing_params is None:
            raise KeyError("waypoint")      # rocket_env.py:168 — the default landing_params lack it
        self.params = derive_params(cfg, {"max_time": 150})
        # the bare env has no TimeLimit / ClipReward (those are make_env() wrappers)
        self._b = Rocket6DOFBatch(1, params=self.params, device=device, seed=seed, auto_reset=False,
                                  clip_reward=False, time_limit=False, debug_buffers=True)
        self.state_names, self.action_names = list(STATE_NAMES), list(ACTION_NAMES)
        self.state_normalizer = self.params.state_normalizer
        self.observation_space = make_box(-1.0, 1.0, (14,), np.float32)
        self.action_space = make_box(-1.0, 1.0, (3,), np.float32)
        self.timestep = timestep
        self.max_thrust, self.max_gimbal = self.params.max_thrust, self.params.max_gimbal
        self.reward_coefficients = self.params.reward_coeff
        self.shaping_type = self.params.shaping_type
        self.state = None
        self.action = np.zeros(3)
        self.state_history, self.action_history, self.timesteps = [], [], []

    def seed(self, seed: int = 42):
        return self._b.seed(seed)

    def reset(self) -> np.ndarray:
        obs = self._b.reset()
        torch.cuda.synchronize(self._b.device)
        self.state = self._b.state[:, 0].cpu().numpy().astype(np.float32)
        self.initial_condition = self.state.copy()
        self.state_history, self.action_history, self.timesteps = [self.state], [[0, 0, 0]], [0]
        return obs[:, 0].cpu().numpy()

    def step(self, normalized_action):
        a = torch.as_tensor(np.asarray(normalized_action, np.float32).reshape(1, 3), device=self._b.device)
        b = self._b
        b.step(a)
        torch.cuda.synchronize(b.device)
        self.state = b.state[:, 0].cpu().numpy()
        fl = int(b.flags[0])
        done = bool(fl & (F_EVENT | F_OOB))
        reward = float(b.reward[0])
        terms = b.reward_terms[:, 0].cpu().numpy()
        rewards_dict = {k: float(v) for k, v in zip(self.params.reward_term_names, terms)}
        gim = np.float32(np.float64(a[0, :2].cpu().numpy()) * self.max_gimbal)
        thr = np.float32((a[0, 2].cpu().numpy() + np.float32(1)) / np.float32(2) * np.float32(self.max_thrust))
        self.action = np.float32([gim[0], gim[1], thr])
        self.state_history.append(self.state)
        self.action_history.append(self.action)
        self.timesteps.append(float(self.params.t_table[min(int(b.step_count[0]), len(self.params.t_table) - 1)]))
        info = {"rewards_dict": rewards_dict, "is_done": done, "state_history": self.state_history,
                "action_history": self.action_history, "timesteps": self.timesteps, **rewards_dict,
                "bounds_violation": bool(fl & F_OOB)}
        return b.obs[:, 0].cpu().numpy(), reward, done, info

    def get_state(self):
        return self.state

    def used_mass(self):
        return self.state_history[0][-1] - self.state_history[-1][-1]

    def _get_normalizer(self):
        return self.state_normalizer

    def render(self, mode: str = "rgb_array"):
        raise NotImplementedError("rendering is stripped from the hot path")

    def close(self) -> None:
        return None

    @property
    def unwrapped(self):
        return self
