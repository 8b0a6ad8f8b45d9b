#!/usr/bin/env python
"""Config 5 at scale: Monte-Carlo landing dispersion of the reference's trained policy over 2^20 episodes
(closed loop fused in the rollout kernel), next to the 30-episode statistics of the reference-side run."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_rocket_6dof_b200 import montecarlo, policy  # noqa: E402

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
gold = os.path.join(root, "tests", "golden", "policy_cl.npz")
w = policy.load_npz(gold)
out = {}
for tk in (False, True):                                   # warm-up: lazy kernel loading, allocator
    montecarlo.run_montecarlo(1 << 20, w, device="cuda:0", seed=1, tensor_cores=True, two_kernel=tk, lanes=2 if tk else 1)
for n, kw in ((30, {}), (1 << 20, dict(two_kernel=False)), (1 << 20, dict(two_kernel=True, lanes=2))):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = montecarlo.run_montecarlo(n, w, device="cuda:0", seed=7, tensor_cores=True, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    key = str(n) + ("" if not kw else "_two_kernel" if kw.get("two_kernel") else "_fused")
    out[key] = dict(seconds=dt, mean=res["mean"], std=res["std"], landed=int(res["landed"].sum()),
                    mean_episode_length=float(res["episode_length"].mean()), env_steps=float(res["stats"]["steps"]),
                    env_steps_per_s=float(res["stats"]["steps"]) / dt, mode=kw)
    print(key, "episodes in %.2f s" % dt, montecarlo.format_report(res), sep="\n")
g = np.load(gold)
ends = [int(s) for s in g["ic_step"]][1:] + [len(g["action"])]
term = np.stack([g["state"][e - 1] for e in ends])
out["reference_30"] = dict(final_position_error=float(np.linalg.norm(term[:, :3], axis=1).mean()),
                           final_velocity_error=float(np.linalg.norm(term[:, 3:6], axis=1).mean()))
json.dump(out, open(sys.argv[1], "w"), indent=1) if len(sys.argv) > 1 else None
