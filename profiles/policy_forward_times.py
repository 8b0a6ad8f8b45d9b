"""Device time of r6_policy_ex (action + value + log-prob) per mode, deterministic and stochastic, 2^20 envs, and of the
per-step trajectory-buffer copies of collect_rollout:   python profiles/policy_forward_times.py   -> one JSON line"""
import json, os, sys
import torch

sys.path.insert(0, os.getcwd())
from rl_rocket_6dof_b200 import policy
from rl_rocket_6dof_b200.batch import Rocket6DOFBatch

n = 1 << 20
env = Rocket6DOFBatch(n, device="cuda:0", seed=42)
env.reset(); env.rollout(64)
w = policy.load_npz("tests/golden/policy_cl.npz")
import numpy as np
w = dict(w)
rs = np.random.default_rng(0)
if "wv" not in w:                               # the fixture holds the actor only: synthetic critic head and log_std
    w["wv"] = (rs.standard_normal(64) * 0.1).astype(np.float32); w["bv"] = np.zeros(1, np.float32)
    w["log_std"] = np.full(3, -0.5, np.float32)
wd = {k: torch.from_numpy(np.ascontiguousarray(v)).to(env.device) for k, v in w.items()}
res = {}
def timeit(f, iters=30):
    for _ in range(3): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return round(ts[len(ts) // 2], 4)
for mode in (3, 2):
    for st in (False, True):
        try:
            res[f"mode{mode}_{'stoch' if st else 'det'}_ms"] = timeit(lambda: env.policy_forward(wd, stochastic=st, tensor_cores=mode, step_index=3))
        except Exception as e:      # weights without a value head
            res["error"] = str(e)[:200]
obs = torch.empty(n, 13, device="cuda"); rew = torch.empty(n, device="cuda"); dn = torch.empty(n, dtype=torch.uint8, device="cuda")
res["obs_copy_ms"] = timeit(lambda: obs.copy_(env.obs[:13].t()))
res["rew_done_copy_ms"] = timeit(lambda: (rew.copy_(env.reward_f32), dn.copy_(env.done)))
print(json.dumps(res))
