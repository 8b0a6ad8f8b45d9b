"""N>1 path on CPU: two `gloo` ranks each own an index range of the environments (stepped with the
TEST-ONLY host build of the kernel code), reduce the 8-slot episode statistics with one all_reduce
and must reproduce the single-process run bit for bit (Philox streams are keyed by global env id)."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

N_GLOBAL, K_STEPS, SEED = 96, 170, 2024


def _run_shard(offset, count, n_global):
    """Reset + K auto-reset steps with Philox actions for envs [offset, offset+count). Returns
    (final state [count,14], stats[8])."""
    import hostsim
    from parity_utils import env_params
    from rl_rocket_6dof_b200._lib import R6Buffers
    ep = env_params()
    L = hostsim.lib()
    hb = hostsim.HostSimBatch(ep, count, auto_reset=True, clip_reward=True, time_limit=True)
    b = R6Buffers()
    b.n_global = n_global
    L.hs_reset(C.byref(hb.p), C.byref(b), hb.envs.ctypes.data, count, offset, SEED)
    stats = np.zeros(8)
    a = (C.c_float * 3)()
    acts = np.zeros((count, 3), np.float32)
    for j in range(K_STEPS):
        for i in range(count):
            L.hs_philox_action(SEED, offset + i, j, a)
            acts[i] = a[:]
        o = hb.step(acts)
        stats[7] += count
        fin = np.nonzero(o["finished"])[0]
        for i in fin:
            fl = int(o["flags"][i])
            stats[0] += 1
            stats[1] += float(hb.envs["ep_return"][i])
            stats[2] += int(hb.envs["k"][i])
            stats[3] += (fl & 0xF8) == 0xF8
            stats[4] += bool(fl & 1)
            stats[5] += bool(fl & 2)
            stats[6] += bool(fl & 4)
        if len(fin):   # auto-reset of the finished envs (episode counter advances the Philox counter)
            sub = hb.envs[fin].copy()
            for r, i in enumerate(fin):
                L.hs_reset(C.byref(hb.p), C.byref(b), sub[r:r + 1].ctypes.data, 1, offset + int(i), SEED)
            hb.envs[fin] = sub
    return hb.envs["y"].copy(), stats


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from rl_rocket_6dof_b200 import sharding
    assert sharding.env_world() == (rank, rank, world)
    assert sharding.init_process_group(torch.device("cpu"))
    off, cnt = sharding.shard_range(N_GLOBAL, world, rank)
    y, stats = _run_shard(off, cnt, N_GLOBAL)
    total = sharding.reduce_stats(torch.from_numpy(stats))
    tmax = sharding.max_over_ranks(float(rank + 1), torch.device("cpu"))
    gathered = [None] * world
    dist.all_gather_object(gathered, (off, cnt, y))
    dist.barrier()
    if rank == 0:
        q.put((total.numpy(), tmax, gathered))
    dist.destroy_process_group()


def test_shard_range_partitions():
    from rl_rocket_6dof_b200.sharding import shard_range
    for n, w in [(0, 1), (1, 1), (7, 2), (96, 2), (1 << 26, 8), (10, 3), (5, 8)]:
        spans = [shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        for (o0, c0), (o1, _) in zip(spans, spans[1:]):
            assert o0 + c0 == o1
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


@pytest.mark.timeout(600)
def test_two_gloo_ranks_match_single_process():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, tmax, gathered = q.get(timeout=500)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    y_ref, stats_ref = _run_shard(0, N_GLOBAL, N_GLOBAL)
    y = np.concatenate([g[2] for g in sorted(gathered, key=lambda g: g[0])])
    assert np.array_equal(y, y_ref)                       # shard-count invariance, bit for bit
    assert np.array_equal(total[[0, 2, 3, 4, 5, 6, 7]], stats_ref[[0, 2, 3, 4, 5, 6, 7]])
    assert abs(total[1] - stats_ref[1]) <= 1e-9 * max(1.0, abs(stats_ref[1]))   # sum order differs
    assert total[7] == N_GLOBAL * K_STEPS and total[0] > 0
    assert tmax == 2.0
