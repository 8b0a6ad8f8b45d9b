#!/usr/bin/env python
"""Per-kernel durations of one env-step (CUPTI via torch.profiler, one stream) + the un-profiled step time on one stream
and on two stream lanes (joined every step / free-running).  A/B tool: R6_LIB_PATH=<variant.so> python profiles/kernel_times.py
    [--envs N] [--steps K] [--tag name] [--no-prof]
Numbers under the profiler are per-kernel SHARES only; the step times printed first are measured without it."""
import argparse
import collections
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_rocket_6dof_b200.batch import Rocket6DOFBatch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--tag", default=os.environ.get("R6_LIB_PATH", "default"))
ap.add_argument("--no-prof", action="store_true")
ap.add_argument("--precision", default="fp64")
ap.add_argument("--preroll", type=int, default=256)
ap.add_argument("--multipass", type=int, default=-1, help="-1 auto, 0 off, 1 on")
a = ap.parse_args()
n, K = a.envs, a.steps
out = {"tag": a.tag, "envs": n}


def timed(env, join, K):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for w in range(4):
        env.step(acts[w % 4], join=join)
    env.join()
    torch.cuda.synchronize()
    e0.record()
    for k in range(K):
        env.step(acts[k % 4], join=join)
    env.join()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


g = torch.Generator(device="cuda"); g.manual_seed(1)
acts = (torch.rand(4, n, 3, device="cuda", generator=g) * 2 - 1).contiguous()
for lanes in (2, 1):
    env = Rocket6DOFBatch(n, device="cuda:0", seed=42, lanes=lanes, record_attempts=True, precision=a.precision,
                          split_step=True, multipass=None if a.multipass < 0 else bool(a.multipass))
    env.reset()
    env.rollout(a.preroll)
    if lanes == 2:
        out["lanes2_joined_ms"] = min(timed(env, True, K) for _ in range(2))
        out["lanes2_free_ms"] = min(timed(env, False, K) for _ in range(2))
    else:
        out["lanes1_ms"] = min(timed(env, True, K) for _ in range(2))
        out["mean_attempts"] = float(env.nattempts.float().mean())
        if not a.no_prof:
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for k in range(10):
                    env.step(acts[k % 4])
                torch.cuda.synchronize()
            import re
            agg = collections.defaultdict(list)
            for ev in prof.events():
                m = re.search(r"(\w+_kernel)", ev.name or "")
                if m and "cuda" in str(ev.device_type).lower():
                    agg[m.group(1)].append(ev.device_time if hasattr(ev, "device_time") else ev.cuda_time)
            # launches of one kernel within a step are listed in launch order (resume pass 0, resume pass 1, ...)
            out["kernels_us"] = {k: [round(sum(v[j::len(v) // 10]) / 10, 2) for j in range(len(v) // 10)] for k, v in agg.items()}
    del env
print(json.dumps(out))
