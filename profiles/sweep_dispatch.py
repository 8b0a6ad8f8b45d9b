#!/usr/bin/env python
"""Where the dispatch thresholds sit with the round-2 kernels: step time vs batch size for the kernel pair / the multi-pass
integrator on one stream / two lanes (joined every step), float64 and float32.  python profiles/sweep_dispatch.py [out.json]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_rocket_6dof_b200.batch import Rocket6DOFBatch
rows = []
for prec in ("fp64", "fp32"):
    for lg in range(15, 22):
        n = 1 << lg
        g = torch.Generator(device="cuda"); g.manual_seed(lg)
        acts = torch.rand(4, n, 3, device="cuda", generator=g) * 2 - 1
        res = {}
        for name, kw in (("fused", dict(split_step=False)), ("pair", dict(split_step=True, multipass=False)),
                         ("pair_l2", dict(split_step=True, multipass=False, lanes=2)),
                         ("multipass", dict(split_step=True, multipass=True)),
                         ("multipass_l2", dict(split_step=True, multipass=True, lanes=2))):
            env = Rocket6DOFBatch(n, device="cuda:0", seed=42, precision=prec, **kw)
            env.reset(); env.rollout(200, fused=True)      # fused pre-roll: the split pre-roll would allocate the pair's scratch
            k = max(20, min(300, (1 << 26) // n))
            for w in range(5): env.step(acts[w % 4])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for j in range(k): env.step(acts[j % 4])
            e1.record(); torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / k * 1e3
            del env
        best = min(res, key=res.get)
        rows.append(dict(precision=prec, envs=n, us_per_step=res, best=best))
        print(prec, f"2^{lg}", {k: round(v, 1) for k, v in res.items()}, "best:", best, flush=True)
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
