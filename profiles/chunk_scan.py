#!/usr/bin/env python
"""Step time vs. (stream lanes, env sub-ranges per step): python profiles/chunk_scan.py [--envs N]"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_rocket_6dof_b200.batch import Rocket6DOFBatch
ap = argparse.ArgumentParser(); ap.add_argument("--envs", type=int, default=1 << 20); ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--combos", default="2x2,2x4,2x8,2x16,3x3,3x6,3x12,4x8,4x16,2x32,3x24")
a = ap.parse_args()
n, K = a.envs, a.steps
g = torch.Generator(device="cuda"); g.manual_seed(1)
acts = (torch.rand(4, n, 3, device="cuda", generator=g) * 2 - 1).contiguous()
for combo in a.combos.split(","):
    lanes, chunks = (int(x) for x in combo.split("x"))
    env = Rocket6DOFBatch(n, device="cuda:0", seed=42, lanes=lanes, chunks=chunks)
    env.reset(); env.rollout(256)
    res = {}
    for join in (True, False):
        for w in range(4): env.step(acts[w % 4], join=join)
        env.join(); torch.cuda.synchronize()
        best = 1e9
        for rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(K): env.step(acts[k % 4], join=join)
            env.join(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / K)
        res["joined" if join else "free"] = best
    print(json.dumps({"lanes": lanes, "chunks": chunks, "joined_us": round(res["joined"] * 1e3, 1), "free_us": round(res["free"] * 1e3, 1)}), flush=True)
    del env
