#!/usr/bin/env python
"""Smallest run that touches every kernel once on a ragged batch of 300 envs, both precisions, fused and split step
(written for compute-sanitizer --tool memcheck, which is closed on this pool; still useful as an all-kernels smoke run)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_rocket_6dof_b200 import policy  # noqa: E402
from rl_rocket_6dof_b200.batch import ACT_BUFFER, ACT_MLP, ACT_MLP_TC, Rocket6DOFBatch  # noqa: E402
from rl_rocket_6dof_b200.gae import compute_gae  # noqa: E402

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
w = policy.load_npz(os.path.join(root, "tests", "golden", "policy_cl.npz"))
w.update(wv=np.full(64, 0.1, np.float32), bv=np.zeros(1, np.float32), log_std=np.full(3, -1.0, np.float32))
n = 300
for prec in ("fp64", "fp32"):
    for split in (False, True):
        env = Rocket6DOFBatch(n, device="cuda:0", seed=1, precision=prec, split_step=split, debug_buffers=True)
        wd = policy.to_device(w, env.device)
        env.reset()
        a = torch.rand(n, 3, device="cuda") * 2 - 1
        for _ in range(3):
            env.step(a)
        env.step_random(2)
        env.rollout(3, fused=True)
        env.rollout(2, ACT_BUFFER, actions=torch.rand(2, n, 3, device="cuda") * 2 - 1, record=True)
        env.rollout(2, ACT_MLP, mlp=wd)
        env.rollout(2, ACT_MLP_TC, mlp=wd, record=True)
        for tc in (0, 1, 2):
            env.policy_forward(wd, stochastic=True, tensor_cores=tc)
        ro = env.collect_rollout(3, wd, tensor_cores=2)
        mask = torch.zeros(n, dtype=torch.uint8, device="cuda"); mask[::3] = 1
        env.reset(mask)
        torch.cuda.synchronize()
compute_gae(torch.rand(4, n, device="cuda"), torch.rand(4, n, device="cuda"), torch.zeros(4, n, dtype=torch.uint8, device="cuda"),
            torch.rand(n, device="cuda"))
torch.cuda.synchronize()
print("sanitize_small: all kernels ran")
