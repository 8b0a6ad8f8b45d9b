import os, sys, time, json
import torch
sys.path.insert(0, os.getcwd())
from rl_rocket_6dof_b200 import montecarlo, policy
w = policy.load_npz("tests/golden/policy_cl.npz")
res = {}
for n in (1 << 12, 1 << 13, 1 << 14, 1 << 15, 1 << 16, 1 << 17):
    for tk in (False, True):
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            montecarlo.run_montecarlo(n, w, device="cuda:0", seed=7, tensor_cores=True, two_kernel=tk, lanes=1)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        res[f"{n}_{'two' if tk else 'fused'}"] = round(dt * 1e3, 1)
print(json.dumps(res))
