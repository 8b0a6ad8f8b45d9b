#!/usr/bin/env python
"""Host-memory ceiling of the host-facing step on N GPUs of one box: every rank copies what one 2^20-env step moves
(58 B/env device->host, 12 B/env host->device, pinned memory, DMA engines, both directions at once) AT THE SAME TIME as
all other ranks, and then runs the real `Rocket6DOFVecEnv.step_host` the same way.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        profiles/pcie_bandwidth_multi.py [--envs 1048576]
Rank 0 prints one JSON line: aggregate copy-engine GB/s (the ceiling), the step's aggregate env-steps/s and the ratio."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--reps", type=int, default=30)
a = ap.parse_args()
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    if world > 1:
        dist.barrier(device_ids=[local])
    torch.cuda.synchronize()


def max_over_ranks(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


n = a.envs
d2h_bytes, h2d_bytes = 58 * n, 12 * n
src_d = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
dst_h = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
src_h = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
dst_d = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
s2 = torch.cuda.Stream()


def both():
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s2):
        dst_d.copy_(src_h, non_blocking=True)
    dst_h.copy_(src_d, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)


def timed(fn, reps):
    for _ in range(3):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return max_over_ranks(e0.elapsed_time(e1) / reps)


ms_copy = timed(both, a.reps)
from rl_rocket_6dof_b200.vec_env import Rocket6DOFVecEnv  # noqa: E402
vec = Rocket6DOFVecEnv(n, device=dev, seed=42, env_offset=rank * n, num_envs_global=world * n)
vec.batch.reset()
vec.batch.rollout(128)
acts = (torch.rand(4, n, 3) * 2 - 1).pin_memory()
k = [0]


def step():
    vec.step_host(acts[k[0] % 4]); k[0] += 1


ms_step = timed(step, a.reps)
if rank == 0:
    print(json.dumps({
        "n_gpus": world, "envs_per_gpu": n, "d2h_bytes_per_gpu": d2h_bytes, "h2d_bytes_per_gpu": h2d_bytes,
        "copy_engines_ms": ms_copy, "copy_engines_aggregate_d2h_GBps": world * d2h_bytes / ms_copy / 1e6,
        "copy_engines_equivalent_env_steps_per_s": world * n / (ms_copy * 1e-3),
        "step_host_ms": ms_step, "step_host_env_steps_per_s": world * n / (ms_step * 1e-3),
        "step_over_copy_ceiling": ms_step / ms_copy,
        "what": "max over ranks; all ranks copy / step concurrently; the copy is a pure transfer of the step's bytes"}), flush=True)
if world > 1:
    dist.destroy_process_group()
