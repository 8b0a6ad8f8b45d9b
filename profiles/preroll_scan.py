#!/usr/bin/env python
"""Step time vs. the number of pre-roll steps: all envs start their first episode together, so the mix of episode phases
(and with it RK attempts per step, resets per step, t_go cold starts) only becomes stationary after several episode
lengths (143 +- 31 steps).  python profiles/preroll_scan.py"""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_rocket_6dof_b200.batch import Rocket6DOFBatch
n = 1 << 20
g = torch.Generator(device="cuda"); g.manual_seed(20)
acts = torch.rand(8, n, 3, device="cuda", generator=g) * 2 - 1
env = Rocket6DOFBatch(n, device="cuda:0", seed=42, lanes=2, record_attempts=True)
env.reset()
done = 0
for pre in (128, 200, 256, 320, 400, 512, 768, 1024, 1536, 2048, 3072):
    env.rollout(pre - done); done = pre
    env.reset_stats()
    for w in range(3): env.step(acts[w % 8])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for j in range(32): env.step(acts[j % 8])
    e1.record(); torch.cuda.synchronize()
    done += 35
    s = env.stats_dict()
    print(json.dumps({"preroll": pre, "us_per_step": round(e0.elapsed_time(e1) / 32 * 1e3, 1), "mean_attempts": round(float(env.nattempts.float().mean()), 3),
                      "episodes_ended_per_step_per_1k_envs": round(s["episodes"] / 35 / n * 1e3, 2)}), flush=True)
