"""Pipe-per-worker multiprocessing VecEnv — TEST / BASELINE INFRASTRUCTURE ONLY.

Stands in for stable_baselines3.common.vec_env.SubprocVecEnv (SB3 is not installed in this image):
one worker process per env, `step` / `reset` / `close` commands over a Pipe, auto-reset on done with
`terminal_observation`, exactly SB3's worker protocol.  Used only to time the CPU arm
(the unmodified reference env from baseline/_ref, or oracle/py_port.py beside it) for bench.py --impl reference and
BASELINE.md's CPU-baseline plan.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _make_port_env(seed):
    from oracle.py_port import WrappedPort
    from rl_rocket_6dof_b200.params import derive_params, load_config
    sb3, cfg = load_config()
    return WrappedPort(derive_params(cfg, sb3), seed)


def _worker(remote, parent_remote, seed, env_fn):
    parent_remote.close()
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    warnings.filterwarnings("ignore")
    env = (env_fn or _make_port_env)(seed)
    try:
        while True:
            cmd, data = remote.recv()
            if cmd == "step":
                obs, r, done, info = env.step(data)
                if done:
                    info["terminal_observation"] = obs
                    obs = env.reset()
                remote.send((obs, r, done, info))
            elif cmd == "reset":
                remote.send(env.reset())
            elif cmd == "close":
                remote.close()
                break
    except (KeyboardInterrupt, EOFError):
        pass


class SubprocVecEnvPort:
    def __init__(self, n_envs, seed0=42, env_fn=None):
        """env_fn(seed) -> gym-style env, called inside each worker (default: the Python/SciPy port with the make_env()
        wrappers; baseline/ref_arm.py passes the factory of the unmodified reference env)."""
        ctx = mp.get_context("fork")
        self.n = n_envs
        self.remotes, work = zip(*[ctx.Pipe() for _ in range(n_envs)])
        self.procs = []
        for i, (w, r) in enumerate(zip(work, self.remotes)):
            p = ctx.Process(target=_worker, args=(w, r, seed0 + i, env_fn), daemon=True)
            p.start()
            self.procs.append(p)
            w.close()

    def reset(self):
        for r in self.remotes:
            r.send(("reset", None))
        return np.stack([r.recv() for r in self.remotes])

    def step(self, actions):
        for r, a in zip(self.remotes, actions):
            r.send(("step", a))
        res = [r.recv() for r in self.remotes]
        obs, rew, done, info = zip(*res)
        return np.stack(obs), np.asarray(rew, np.float32), np.asarray(done), list(info)

    def close(self):
        for r in self.remotes:
            try:
                r.send(("close", None))
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=5)


def time_python_port(n_workers, steps, warmup):
    """env-steps/s of the Python/SciPy port under the SubprocVecEnv-style harness."""
    vec = SubprocVecEnvPort(n_workers)
    try:
        vec.reset()
        rngs = [np.random.default_rng(i) for i in range(n_workers)]

        def acts():
            return [r.uniform(-1, 1, 3).astype(np.float32) for r in rngs]
        for _ in range(warmup):
            vec.step(acts())
        t0 = time.perf_counter()
        for _ in range(steps):
            vec.step(acts())
        dt = time.perf_counter() - t0
    finally:
        vec.close()
    return {"steps_per_s": n_workers * steps / dt, "seconds": dt, "vec_steps": steps,
            "sample": f"oracle/py_port.py (Python + SciPy solve_ivp restatement of the reference env, make_env() "
                      f"wrappers), {n_workers} worker processes x {steps} steps, pipe-based SubprocVecEnv protocol"}
