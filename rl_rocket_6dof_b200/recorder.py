"""Trajectory recorder for the batched env: the `[T, N, ...]` equivalent of the reference simulator's per-episode lists
`SIM.states` / `SIM.actions` / `SIM.times` (simulator.py:30-35,100-102), which `info["state_history"]`,
`info["action_history"]`, `info["timesteps"]` (rocket_env.py:219-223) and the dataframe getters
(`states_to_dataframe` / `actions_to_dataframe`, rocket_env.py:572-579) expose.

The step kernels keep no history (the hot path writes only the current state), so recording is opt-in and lives on
the host side of the boundary: after `reset()` and after every `step(actions)` the recorder copies the raw float64
state `[14, N]` and the denormalised action into preallocated device tensors — two device-to-device copies per step
on the env's stream, no synchronisation.  Row 0 is the reset row, as in the reference: state = initial condition,
action = [0, 0, 0], time = 0; rows 1..length are the stepped rows.  With auto-reset on, the row written for a step
that ended an episode holds the TERMINAL state (what the reference appended before `done`), the initial condition
the kernel drew for the next episode goes to `initial[t+1]` with `episode_start[t+1, i]` set, and the step clock of
that env restarts; `episode(i, j)` cuts one env's j-th episode out of the record in the reference's list shape.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .batch import Rocket6DOFBatch

STATE_NAMES = ["x", "y", "z", "vx", "vy", "vz", "q0", "q1", "q2", "q3", "omega1", "omega2", "omega3", "mass"]
ACTION_NAMES = ["gimbal_y", "gimbal_z", "thrust"]


class TrajectoryRecorder:
    def __init__(self, batch: Rocket6DOFBatch, capacity: int):
        if capacity <= 0:
            raise ValueError("capacity must be positive")
        self.batch, self.capacity = batch, int(capacity)
        n, dev = batch.num_envs, batch.device
        sdt = batch.state.dtype
        self.states = torch.empty(self.capacity + 1, 14, n, dtype=sdt, device=dev)
        self.actions = torch.zeros(self.capacity + 1, n, 3, dtype=torch.float32, device=dev)
        self.times = torch.zeros(self.capacity + 1, n, dtype=torch.float64, device=dev)
        # episode_start[t, i]: the episode that stepped row t of env i belongs to begins there, from initial[t, :, i]
        self.episode_start = torch.zeros(self.capacity + 2, n, dtype=torch.bool, device=dev)
        self.initial = torch.empty(self.capacity + 2, 14, n, dtype=sdt, device=dev)
        self._t_table = torch.as_tensor(np.asarray(batch.params.t_table, np.float64), device=dev)
        self._clock = torch.zeros(n, dtype=torch.int64, device=dev)
        self.length = 0                               # rows written after the reset row

    # ------------------------------------------------------------------ recording
    def reset(self):
        """`batch.reset()` + row 0 of the record."""
        obs = self.batch.reset()
        self.length = 0
        self.states[0].copy_(self.batch.state)
        self.actions[0].zero_()
        self.times[0].zero_()
        self.episode_start.zero_()
        self.episode_start[1].fill_(True)
        self.initial[1].copy_(self.batch.state)
        self._clock.zero_()
        return obs

    def denormalize(self, actions: torch.Tensor) -> torch.Tensor:
        """`_denormalize_action` (rocket_env.py:509-521) with its precision map: gimbals through the float64
        `max_gimbal`, thrust in float32."""
        p = self.batch.params
        u = torch.empty_like(actions)
        u[:, :2] = (actions[:, :2].double() * float(p.max_gimbal)).float()
        u[:, 2] = (actions[:, 2] + 1.0) / 2.0 * float(np.float32(p.max_thrust))
        return u

    def step(self, actions: torch.Tensor):
        """`batch.step(actions)` + one row of the record."""
        if self.length >= self.capacity:
            raise RuntimeError("trajectory record is full")
        b = self.batch
        out = b.step(actions)
        t = self.length + 1
        done = b.done.bool()
        # at done the live state is already the next episode's initial condition (auto-reset): the record takes
        # the terminal state the kernel parked in `terminal_state`
        if b.auto_reset:
            self.states[t].copy_(torch.where(done.unsqueeze(0), b.terminal_state.to(self.states.dtype), b.state))
        else:
            self.states[t].copy_(b.state)
        self.actions[t].copy_(self.denormalize(actions))
        self._clock += 1
        self.times[t].copy_(self._t_table[self._clock.clamp_(max=self._t_table.numel() - 1)])
        if b.auto_reset:
            self.episode_start[t + 1].copy_(done)
            self.initial[t + 1].copy_(b.state)
            self._clock.masked_fill_(done, 0)
        self.length = t
        return out

    # ------------------------------------------------------------------ read-out
    def episode(self, env: int, index: int = 0) -> dict:
        """The `index`-th episode of env `env` as numpy arrays shaped like the reference's lists: states [L+1, 14],
        actions [L+1, 3], times [L+1], row 0 = (initial condition, [0, 0, 0], 0).  An episode cut by the end of
        the record is returned as far as it got."""
        last = self.length + 1
        starts = (torch.nonzero(self.episode_start[1:last + 1, env]).flatten() + 1).tolist()
        if index >= len(starts):
            raise IndexError("episode index beyond the record")
        lo = starts[index]
        hi = starts[index + 1] if index + 1 < len(starts) else last
        st = torch.cat([self.initial[lo, :, env].unsqueeze(0), self.states[lo:hi, :, env]])
        ac = torch.cat([torch.zeros(1, 3, dtype=torch.float32, device=st.device), self.actions[lo:hi, env]])
        tm = torch.cat([torch.zeros(1, dtype=torch.float64, device=st.device), self.times[lo:hi, env]])
        return {"states": st.cpu().numpy(), "actions": ac.cpu().numpy(), "times": tm.cpu().numpy()}

    def states_to_dataframe(self, env: int = 0, index: int = 0):
        import pandas as pd
        return pd.DataFrame(self.episode(env, index)["states"], columns=STATE_NAMES)

    def actions_to_dataframe(self, env: int = 0, index: int = 0):
        import pandas as pd
        return pd.DataFrame(self.episode(env, index)["actions"], columns=ACTION_NAMES)

    def used_mass(self, env: int = 0, index: int = 0) -> float:
        """rocket_env.py:585-589."""
        s = self.episode(env, index)["states"]
        return float(s[0, -1] - s[-1, -1])


def record_rollout(batch: Rocket6DOFBatch, k: int, actions: Optional[torch.Tensor] = None, mlp: Optional[dict] = None,
                   tensor_cores=False) -> TrajectoryRecorder:
    """Resets the batch and records k steps driven by a `[k, N, 3]` action tensor or by the deterministic policy."""
    if (actions is None) == (mlp is None):
        raise ValueError("give either actions [k, N, 3] or the policy weights")
    rec = TrajectoryRecorder(batch, k)
    rec.reset()
    for j in range(int(k)):
        a = actions[j] if actions is not None else batch.policy_actions(mlp, tensor_cores=tensor_cores)
        rec.step(a)
    return rec
