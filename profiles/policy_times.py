"""Device time of the policy kernels (r6_policy, tensor_cores = 0..3) on 2^20 envs and their action error against the
float32 FMA network:   R6_LIB_PATH=<lib> python profiles/policy_times.py [--tag T]     -> one JSON line"""
import argparse, json, os, sys
import torch

sys.path.insert(0, os.getcwd())
from rl_rocket_6dof_b200 import policy
from rl_rocket_6dof_b200.batch import Rocket6DOFBatch

ap = argparse.ArgumentParser()
ap.add_argument("--tag", default="")
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--iters", type=int, default=50)
a = ap.parse_args()
n = a.envs
env = Rocket6DOFBatch(n, device="cuda:0", seed=42)
env.reset(); env.rollout(64)
w = policy.to_device(policy.load_npz("tests/golden/policy_cl.npz"), env.device)
ref = torch.empty(n, 3, device="cuda")
env.policy_actions(w, tensor_cores=0, out=ref)
res = {"tag": a.tag, "envs": n}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for mode in (0, 1, 2, 3):
    out = torch.empty(n, 3, device="cuda")
    for _ in range(3):
        env.policy_actions(w, tensor_cores=mode, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); env.policy_actions(w, tensor_cores=mode, out=out); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    res[f"mode{mode}_ms"] = round(ts[len(ts) // 2], 4)
    res[f"mode{mode}_maxerr"] = float((out - ref).abs().max())
print(json.dumps(res))
