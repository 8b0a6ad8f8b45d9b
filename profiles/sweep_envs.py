#!/usr/bin/env python
"""Config 3 of BASELINE.json: random-action throughput sweep over the batch size on one B200, float64 and
float32 paths (one r6_step launch per env-step, actions resident in HBM, auto-reset on, 256 pre-roll steps).
    python profiles/sweep_envs.py [out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_rocket_6dof_b200.batch import Rocket6DOFBatch  # noqa: E402

rows = []
for prec in ("fp64", "fp32"):
  for lanes in (1, 2):
    for lg in range(14 if lanes == 1 else 17, 24):
        n = 1 << lg
        env = Rocket6DOFBatch(n, device="cuda:0", seed=42, precision=prec, record_attempts=True, lanes=lanes)
        env.reset()
        env.rollout(256)
        g = torch.Generator(device="cuda"); g.manual_seed(lg)
        acts = torch.rand(4, n, 3, device="cuda", generator=g) * 2 - 1
        k = max(20, min(400, (1 << 27) // n))
        for w in range(5):
            env.step(acts[w % 4], join=False)
        env.join()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for j in range(k):
            env.step(acts[j % 4], join=False)           # lanes run free, joined once before the closing event
        env.join()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / k
        att = float(env.nattempts.float().mean())
        rows.append(dict(precision=prec, envs=n, lanes=lanes, multipass=env.work is not None, ms_per_step=ms, env_steps_per_s=n / ms * 1e3, mean_rk_attempts=att,
                         steps_timed=k))
        print(f"{prec} lanes {lanes} 2^{lg:2d} = {n:8d} envs: {ms * 1e3:9.1f} us/step  {n / ms * 1e3:.3e} env-steps/s  (attempts {att:.3f})",
              flush=True)
        del env
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
