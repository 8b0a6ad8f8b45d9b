"""Rocket6DOFVecEnv — stable-baselines3-compatible VecEnv over the batched CUDA env.

Drop-in for what SB3 builds around the reference's `make_env()` (main_6DOF.py:44-53):
`Monitor(TimeLimit(ClipReward(RemoveMassFromObs(Rocket6DOF))))` inside a `DummyVecEnv`, i.e.
13-dim float32 observations, rewards clipped to [-1, 100] as float32, `dones`, auto-reset with
`infos[i]["terminal_observation"]`, Monitor's `infos[i]["episode"] = {"r", "l", "t"}` and
`infos[i]["TimeLimit.truncated"]`; at done the info also carries the terminal 14-state under
`state_history[-1]`, which is what montecarlo_script.py:33-40 reads.

Two call levels:
  * `step_host(actions)` / `reset_host()`: pinned-host-buffer fast path (no Python per-env work);
    this is what scales to 2^20 envs and what bench.py's e2e number measures.  With `zero_copy=True`
    (default) the step kernel reads the actions from, and writes obs / reward / done / flags to,
    pinned device-mapped host memory directly, so the PCIe traffic overlaps the computation instead
    of being three serial copies around it;
  * `reset()`, `step_async()`, `step_wait()`, `step()`, ...: the SB3 VecEnv protocol on top of it
    (info dicts are materialised only for envs that finished).
If stable-baselines3 is importable the class is registered as a virtual subclass of its VecEnv.
"""
from __future__ import annotations

import time
from typing import Any, List, Optional, Sequence

import ctypes as C

import numpy as np
import torch

from . import _lib
from .batch import F_EVENT, F_OOB, F_TRUNCATED, F_LANDING_ALL, FLAG_NAMES, Rocket6DOFBatch
from .spaces import make_box


def _flag_info(f: int) -> dict:
    """The info entries that depend only on the step's flag byte (R6_F_*)."""
    return {"TimeLimit.truncated": bool(f & F_TRUNCATED), "is_done": bool(f & (F_EVENT | F_OOB)),
            "bounds_violation": bool(f & F_OOB),
            "landing_conditions": {nm: bool(f & (8 << k)) for k, nm in enumerate(FLAG_NAMES)},
            "is_successful": (f & F_LANDING_ALL) == F_LANDING_ALL}


_FLAG_INFO = [_flag_info(f) for f in range(256)]       # templates: copied per finished env, never handed out


class MonitorTag:
    """Stands for stable_baselines3.common.monitor.Monitor in `env_is_wrapped` queries."""


class Rocket6DOFVecEnv:
    metadata = {"render.modes": []}

    def __init__(self, num_envs: int, env_config: Optional[dict] = None, sb3_config: Optional[dict] = None, *,
                 device="cuda", seed: Optional[int] = None, remove_mass_from_obs: bool = True,
                 clip_reward: bool = True, time_limit: bool = True, zero_copy: bool = True, **batch_kw):
        self.batch = Rocket6DOFBatch(num_envs, env_config, sb3_config, device=device, seed=seed, auto_reset=True,
                                     clip_reward=clip_reward, time_limit=time_limit, **batch_kw)
        self.num_envs = int(num_envs)
        self.obs_dim = 13 if remove_mass_from_obs else 14
        self.observation_space = make_box(-1.0, 1.0, (self.obs_dim,), np.float32)
        self.action_space = make_box(-1.0, 1.0, (3,), np.float32)
        self.reward_range = (self.batch.params.clip_lo, self.batch.params.clip_hi) if clip_reward else (-np.inf, np.inf)
        n, dev = self.num_envs, self.batch.device
        # pinned staging buffers (host side of the boundary)
        self._act_h = torch.empty(n, 3, dtype=torch.float32).pin_memory()
        self._act_d = torch.empty(n, 3, dtype=torch.float32, device=dev)
        self._obs_h = torch.empty(self.obs_dim, n, dtype=torch.float32).pin_memory()
        self._rew_h = torch.empty(n, dtype=torch.float32).pin_memory()
        self._done_h = torch.empty(n, dtype=torch.uint8).pin_memory()
        self._flags_h = torch.empty(n, dtype=torch.uint8).pin_memory()
        self._actions = None
        self._act_in_flight = None
        self._t_start = time.time()
        self.zero_copy = bool(zero_copy)
        if self.zero_copy:
            # same device state, but the per-step outputs land in the pinned host buffers (UVA: a pinned
            # torch tensor's data_ptr() is valid on the device) and only `obs_dim` observation rows are written
            b = self.batch
            # the kernel writes the observations row-major [N, obs_dim] — the array the VecEnv protocol returns
            self._p_host = b.params.to_struct(**{**b._struct_kw, "obs_rows": self.obs_dim, "obs_row_major": True})
            self._obs_rm = torch.empty(n, self.obs_dim, dtype=torch.float32).pin_memory()
            hb = type(b._b)()
            for name, _ in hb._fields_:
                setattr(hb, name, getattr(b._b, name))
            hb.obs, hb.reward, hb.reward_f32 = self._obs_rm.data_ptr(), 0, self._rew_h.data_ptr()
            hb.done, hb.flags = self._done_h.data_ptr(), self._flags_h.data_ptr()
            # one fused kernel here, not the integrator | post-step pair: its PCIe writes then overlap the
            # integration of other warps (1.27 ms per 2^20-env step vs 1.64 ms with the split, which would also
            # read the host-resident actions twice)
            hb.scratch = 0
            self._b_host = hb
        self.h2d_bytes_per_step = self._act_h.numel() * 4
        self.d2h_bytes_per_step = self._obs_h.numel() * 4 + self._rew_h.numel() * 4 + 2 * n

    # ------------------------------------------------------------------ fast path
    def reset_host(self) -> np.ndarray:
        b = self.batch
        b.reset()
        self._obs_h.copy_(b.obs[: self.obs_dim], non_blocking=True)
        torch.cuda.current_stream(b.device).synchronize()
        return self._obs_h.numpy().T          # [N, obs_dim] view (strided)

    def step_host(self, actions) -> tuple:
        """actions: [N,3] float32 (numpy or CPU tensor).  Returns numpy views of the pinned buffers (valid until
        the next call): obs [N, obs_dim], rewards [N] f32, dones [N] bool.  zero_copy=True: one kernel that reads the
        (pinned) actions and writes obs / reward / done / flags straight to host memory over PCIe — the device-side
        `batch.obs` is then NOT refreshed (use `batch.step` / `batch.policy_*` for device-resident loops).
        zero_copy=False: host->device copy, step kernels, device->host copies."""
        self._launch_host_step(actions)
        return self._finish_host_step()

    def _launch_host_step(self, actions) -> None:
        """Everything of `step_host` up to (not including) the wait for the device: returns once the work is enqueued."""
        b = self.batch
        a = torch.as_tensor(actions, dtype=torch.float32).reshape(self.num_envs, 3)
        if self.zero_copy:
            b.join()
            if not (a.is_pinned() and a.is_contiguous()):
                self._act_h.copy_(a)
                a = self._act_h
            self._act_in_flight = a                 # keep the pinned source alive until the kernel has read it
            with torch.cuda.device(b.device):
                _lib.check(b.lib.r6_step(C.byref(self._p_host), C.byref(self._b_host), self.num_envs, b.env_offset,
                                         a.data_ptr(), b.seed_value, b._stream()), b.lib)
            b.steps_done += 1
            return
        if a.is_pinned():                      # caller already staged the actions in pinned memory
            self._act_in_flight = a
            self._act_d.copy_(a, non_blocking=True)
        else:
            self._act_h.copy_(a)
            self._act_d.copy_(self._act_h, non_blocking=True)
        b.step(self._act_d)
        self._obs_h.copy_(b.obs[: self.obs_dim], non_blocking=True)
        self._rew_h.copy_(b.reward_f32, non_blocking=True)
        self._done_h.copy_(b.done, non_blocking=True)
        self._flags_h.copy_(b.flags, non_blocking=True)

    def _finish_host_step(self) -> tuple:
        b = self.batch
        torch.cuda.current_stream(b.device).synchronize()
        self._act_in_flight = None
        if self.zero_copy:
            return self._obs_rm.numpy(), self._rew_h.numpy(), self._done_h.numpy().view(np.bool_)
        return self._obs_h.numpy().T, self._rew_h.numpy(), self._done_h.numpy().view(np.bool_)

    # ------------------------------------------------------------------ SB3 VecEnv protocol
    def reset(self) -> np.ndarray:
        return np.ascontiguousarray(self.reset_host())

    def step_async(self, actions) -> None:
        """Enqueues the step and returns (the SubprocVecEnv contract: the caller may work until `step_wait`)."""
        self._launch_host_step(np.asarray(actions, dtype=np.float32))
        self._actions = True

    def step_wait(self):
        if self._actions is None:
            raise RuntimeError("step_wait without step_async")
        self._actions = None
        obs, rews, dones = self._finish_host_step()
        # [N, obs_dim] contiguous: torch's blocked transpose-copy of the pinned [obs_dim, N] buffer is several times
        # faster than numpy's strided copy
        obs = obs.copy() if obs.flags["C_CONTIGUOUS"] else self._obs_h.t().contiguous().numpy()
        # info dicts are only filled in for envs that finished; every other env gets a FRESH empty dict each step, so a
        # wrapper or callback that writes into infos[i] cannot leak keys into later steps
        infos: List[dict] = [{} for _ in range(self.num_envs)]
        idx = np.nonzero(dones)[0]
        if len(idx):
            b = self.batch
            sel = torch.as_tensor(idx, device=b.device)
            tobs = b.terminal_obs[: self.obs_dim][:, sel].t().cpu().numpy()
            tstate = b.terminal_state[:, sel].t().cpu().numpy()
            epi = b.ep_info[:, sel].cpu().numpy()
            fl = self._flags_h.numpy()[idx].tolist()
            ret, length = epi[0].tolist(), epi[1].astype(np.int64).tolist()
            now = round(time.time() - self._t_start, 6)
            for j, i in enumerate(idx.tolist()):
                tpl = _FLAG_INFO[fl[j]]
                d = dict(tpl)
                d["landing_conditions"] = dict(tpl["landing_conditions"])
                d["terminal_observation"] = tobs[j]
                d["episode"] = {"r": round(ret[j], 6), "l": length[j], "t": now}      # Monitor: float64 sum, 6 decimals
                d["state_history"] = [tstate[j]]
                infos[i] = d
        return obs, rews.copy(), dones.copy(), infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self) -> None:
        return None

    def seed(self, seed: Optional[int] = None) -> List[Optional[int]]:
        if seed is not None:
            self.batch.seed(seed)
        return [None if seed is None else seed + i for i in range(self.num_envs)]

    def render(self, mode: str = "human"):
        raise NotImplementedError("rendering is stripped from the hot path (pyvista scene of the reference)")

    def get_images(self) -> Sequence[np.ndarray]:
        return []

    def _indices(self, indices):
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, int):
            return [indices]
        return indices

    def get_attr(self, attr_name: str, indices=None) -> List[Any]:
        p = self.batch.params
        table = {"state_normalizer": p.state_normalizer, "max_thrust": p.max_thrust, "max_gimbal": p.max_gimbal,
                 "timestep": p.timestep, "reward_coefficients": p.reward_coeff, "shaping_type": p.shaping_type,
                 "observation_space": self.observation_space, "action_space": self.action_space,
                 "reward_range": self.reward_range, "spec": None, "render_mode": None}
        if attr_name not in table:
            raise AttributeError(attr_name)
        return [table[attr_name] for _ in self._indices(indices)]

    def set_attr(self, attr_name: str, value: Any, indices=None) -> None:
        raise AttributeError(f"{attr_name} is fixed at construction (kernel parameters)")

    def env_method(self, method_name: str, *args, indices=None, **kwargs) -> List[Any]:
        if method_name == "get_state":
            st = self.batch.state.t().cpu().numpy()
            return [st[i] for i in self._indices(indices)]
        if method_name == "seed":
            return self.seed(*args, **kwargs)
        raise AttributeError(method_name)

    def env_is_wrapped(self, wrapper_class, indices=None) -> List[bool]:
        name = getattr(wrapper_class, "__name__", "")
        ok = name in ("Monitor", "MonitorTag", "TimeLimit", "ClipReward", "RemoveMassFromObs")
        return [ok for _ in self._indices(indices)]

    @property
    def unwrapped(self):
        return self

    def getattr_depth_check(self, name: str, already_found: bool):
        return None


try:  # register with SB3 when it exists so isinstance(env, VecEnv) holds
    from stable_baselines3.common.vec_env import VecEnv as _SB3VecEnv  # type: ignore

    _SB3VecEnv.register(Rocket6DOFVecEnv)
except Exception:  # pragma: no cover - SB3 is not installed in this image
    pass
