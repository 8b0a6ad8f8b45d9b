"""gym 0.21 `gym.wrappers.TimeLimit` stand-in (test / baseline infrastructure only, see gym/__init__.py).
Semantics [memory, gym 0.21 wrappers/time_limit.py]: count steps since reset; at max_episode_steps set
info["TimeLimit.truncated"] = not done and done = True."""
from . import Wrapper


class TimeLimit(Wrapper):
    def __init__(self, env, max_episode_steps=None):
        super().__init__(env)
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = None

    def step(self, action):
        assert self._elapsed_steps is not None, "Cannot call env.step() before calling reset()"
        observation, reward, done, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            info["TimeLimit.truncated"] = not done
            done = True
        return observation, reward, done, info

    def reset(self, **kwargs):
        self._elapsed_steps = 0
        return self.env.reset(**kwargs)
