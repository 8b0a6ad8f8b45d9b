#!/usr/bin/env python
"""How well does an env's RK attempt count at step k-1 predict the one at step k, and what would regrouping the envs
of a launch by that prediction buy?  Warp cost model: a warp runs max(attempts) of its 32 lanes.
Prints, for the bench workload (2^20 envs, random actions, steady-state episode mix): P(n_k == n_{k-1}), the mean
attempts per lane, and attempt-warps per warp for (a) env-id order, (b) envs sorted by n_{k-1}, (c) sorted by n_k
itself (oracle grouping)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from rl_rocket_6dof_b200.batch import Rocket6DOFBatch  # noqa: E402

N, K = 1 << 20, 12


def warp_cost(n_att, order=None):
    x = n_att if order is None else n_att[order]
    return float(x.view(-1, 32).max(dim=1).values.double().mean())


def main():
    env = Rocket6DOFBatch(N, seed=42, record_attempts=True)
    env.reset()
    env.rollout(256)
    gen = torch.Generator(device="cuda"); gen.manual_seed(1)
    prev = None
    rows = []
    for k in range(K):
        a = torch.rand(N, 3, device="cuda", generator=gen) * 2 - 1
        env.step(a)
        cur = env.nattempts.clone().long()
        if prev is not None:
            same = float((cur == prev).double().mean())
            order_prev = torch.argsort(prev, stable=True)
            order_cur = torch.argsort(cur, stable=True)
            rows.append({"p_same": same, "mean_attempts": float(cur.double().mean()),
                         "warp_cost_env_order": warp_cost(cur), "warp_cost_sorted_by_prev": warp_cost(cur, order_prev),
                         "warp_cost_sorted_by_self": warp_cost(cur, order_cur),
                         "hist": torch.bincount(cur, minlength=6)[:6].tolist()})
        prev = cur
    keys = ("p_same", "mean_attempts", "warp_cost_env_order", "warp_cost_sorted_by_prev", "warp_cost_sorted_by_self")
    print(json.dumps({"mean": {k: sum(r[k] for r in rows) / len(rows) for k in keys}, "last": rows[-1]}))


if __name__ == "__main__":
    main()
