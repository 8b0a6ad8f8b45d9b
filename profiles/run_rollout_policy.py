#!/usr/bin/env python
"""Small driver for profiling the closed-loop rollout kernels: python profiles/run_rollout_policy.py [tc|cc] [envs] [k]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_rocket_6dof_b200 import policy  # noqa: E402
from rl_rocket_6dof_b200.batch import ACT_MLP, ACT_MLP_TC, Rocket6DOFBatch  # noqa: E402

mode = ACT_MLP_TC if (len(sys.argv) < 2 or sys.argv[1] == "tc") else ACT_MLP
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
k = int(sys.argv[3]) if len(sys.argv) > 3 else 16
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
w = policy.to_device(policy.load_npz(os.path.join(root, "tests", "golden", "policy_cl.npz")), "cuda:0")
env = Rocket6DOFBatch(n, device="cuda:0", seed=42)
env.reset()
env.rollout(64)
for _ in range(2):
    env.rollout(k, mode, mlp=w)
torch.cuda.synchronize()
t0 = time.perf_counter()
env.rollout(k, mode, mlp=w)
torch.cuda.synchronize()
print(f"mode {mode}: {1e3 * (time.perf_counter() - t0) / k:.4f} ms per step of {n} envs")
