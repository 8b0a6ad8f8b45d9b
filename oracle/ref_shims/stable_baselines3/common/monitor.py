"""stable_baselines3.common.monitor.Monitor stand-in (SB3 is not installed in this image) — test / baseline
infrastructure only.  Semantics [memory, SB3 1.6.0 common/monitor.py]: collect the rewards of the running episode; at
done put info["episode"] = {"r": round(sum, 6), "l": length, "t": round(seconds since creation, 6)}."""
import time

import gym


class Monitor(gym.Wrapper):
    def __init__(self, env, filename=None, allow_early_resets=True, reset_keywords=(), info_keywords=()):
        super().__init__(env)
        self.t_start = time.time()
        self.rewards = None
        self.needs_reset = True
        self.episode_returns, self.episode_lengths, self.episode_times = [], [], []
        self.total_steps = 0

    def reset(self, **kwargs):
        self.rewards = []
        self.needs_reset = False
        return self.env.reset(**kwargs)

    def step(self, action):
        if self.needs_reset:
            raise RuntimeError("Tried to step environment that needs reset")
        observation, reward, done, info = self.env.step(action)
        self.rewards.append(reward)
        if done:
            self.needs_reset = True
            ep_rew, ep_len = sum(self.rewards), len(self.rewards)
            ep_info = {"r": round(ep_rew, 6), "l": ep_len, "t": round(time.time() - self.t_start, 6)}
            self.episode_returns.append(ep_rew)
            self.episode_lengths.append(ep_len)
            self.episode_times.append(time.time() - self.t_start)
            info["episode"] = ep_info
        self.total_steps += 1
        return observation, reward, done, info
