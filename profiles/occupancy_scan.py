"""Step-kernel time vs. number of resident warps per SM sub-partition (latency-vs-throughput scan).
Run on the GPU box:  R6_LIB_PATH=<variant.so> python profiles/occupancy_scan.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_rocket_6dof_b200.batch import Rocket6DOFBatch

res = {}
for w in (1, 2, 3, 4, 6, 12, 24, 48):
    n = 148 * 4 * 32 * w
    env = Rocket6DOFBatch(n, seed=1)
    env.reset(); env.rollout(200)
    acts = (torch.rand(4, n, 3, device="cuda") * 2 - 1).contiguous()
    for k in range(3):
        env.step(acts[k % 4])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    K = 20
    for k in range(K):
        env.step(acts[k % 4])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    res[w] = ms
    print(f"warps per SMSP offered = {w:3d}  envs = {n:8d}  step = {ms*1e3:8.1f} us  -> {n/ms/1e6:8.1f} M env-steps/s", flush=True)
print(json.dumps(res))
