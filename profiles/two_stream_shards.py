#!/usr/bin/env python
"""Experiment: does running the 2^20-env step as S independent env-range shards on S CUDA streams of ONE GPU hide the
tail of each kernel (integrate: 18.45 waves of one-warp CTAs; post: similar) behind the other shards' work?
Shards are what multi-GPU sharding already uses (Philox keyed by global env id => same trajectories).
Prints ms per 2^20-env step for S = 1, 2, 4 (interleaved repetitions on the same box)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from rl_rocket_6dof_b200.batch import Rocket6DOFBatch  # noqa: E402
from rl_rocket_6dof_b200.sharding import shard_range  # noqa: E402

N = 1 << 20
K, W, REPS = 50, 100, 5


def make(s):
    out = []
    for r in range(s):
        off, cnt = shard_range(N, s, r)
        b = Rocket6DOFBatch(cnt, seed=42, env_offset=off, num_envs_global=N, split_step=True)
        b.reset()
        out.append((b, torch.cuda.Stream()))
    return out


def run(shards, k, join_every_step=False):
    main = torch.cuda.current_stream()
    if join_every_step:                 # what a stream-ordered r6_step would have to do: fork and join inside each call
        for _ in range(k):
            for b, st in shards:
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    b.step_random(1)
            for b, st in shards:
                main.wait_stream(st)
        return
    for b, st in shards:
        st.wait_stream(main)
    for _ in range(k):
        for b, st in shards:
            with torch.cuda.stream(st):
                b.step_random(1)
    for b, st in shards:
        main.wait_stream(st)


def main():
    sets = {s: make(s) for s in (1, 2, 4)}
    for s in sets:
        run(sets[s], W)
    torch.cuda.synchronize()
    cases = [(s, j) for s in sets for j in ((False, True) if s > 1 else (False,))]
    res = {c: [] for c in cases}
    for _ in range(REPS):
        for s, j in cases:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            run(sets[s], K, j)
            e1.record()
            torch.cuda.synchronize()
            res[(s, j)].append(e0.elapsed_time(e1) / K)
    out = {f"streams_{s[0]}{'_join_every_step' if s[1] else ''}": {"ms_per_step": sorted(v)[len(v) // 2], "all": [round(x, 4) for x in v]} for s, v in res.items()}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
