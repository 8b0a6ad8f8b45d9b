"""NumPy restatement of stable-baselines3 1.6.0 `RolloutBuffer.compute_returns_and_advantage`
(stable_baselines3/common/buffers.py) — TEST INFRASTRUCTURE ONLY.

SB3 is a third-party dependency of the reference (requirements: stable-baselines3 1.6.0; not vendored under
/root/reference and not installed in this image), so this follows its published algorithm: float32 buffers,
python-float gamma / gae_lambda, backward recursion with `next_non_terminal = 1 - episode_starts[t+1]` and, for
the last step, `1 - dones`.  With `dones[t]` = "episode ended at step t", `episode_starts[t+1] == dones[t]`.
Parity against SB3 itself is unpinned (it cannot be run here); the tests pin the kernel to this statement and
check the closed forms of the recursion.
"""
import numpy as np


def compute_returns_and_advantage(rewards, values, dones, last_values, gamma=0.99, gae_lambda=0.95):
    rewards = np.asarray(rewards, np.float32)
    values = np.asarray(values, np.float32)
    dones = np.asarray(dones).astype(np.float32)
    last_values = np.asarray(last_values, np.float32)
    T = rewards.shape[0]
    adv = np.zeros_like(rewards)
    last_gae_lam = 0
    for step in reversed(range(T)):
        next_non_terminal = np.float32(1.0) - dones[step]
        next_values = last_values if step == T - 1 else values[step + 1]
        delta = rewards[step] + gamma * next_values * next_non_terminal - values[step]
        last_gae_lam = delta + gamma * gae_lambda * next_non_terminal * last_gae_lam
        adv[step] = last_gae_lam
    return adv, adv + values
