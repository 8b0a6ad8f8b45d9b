#!/usr/bin/env python
"""Driver for profiling r6_policy alone: python profiles/run_policy_kernel.py [cc|tc|tcgen05] [envs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_rocket_6dof_b200 import policy  # noqa: E402
from rl_rocket_6dof_b200.batch import Rocket6DOFBatch  # noqa: E402

tc = {"cc": 0, "tc": 1, "tcgen05": 2}[sys.argv[1] if len(sys.argv) > 1 else "tc"]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
w = policy.to_device(policy.load_npz(os.path.join(root, "tests", "golden", "policy_cl.npz")), "cuda:0")
env = Rocket6DOFBatch(n, device="cuda:0", seed=42)
env.reset()
env.rollout(32)
for _ in range(3):
    env.policy_actions(w, tensor_cores=tc)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    env.policy_actions(w, tensor_cores=tc)
e1.record()
torch.cuda.synchronize()
print(f"r6_policy tensor_cores={tc}: {e0.elapsed_time(e1) / 20:.4f} ms for {n} envs")
