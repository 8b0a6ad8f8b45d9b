#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE — test infrastructure only.

Run in the build container (needs /root/reference):

    python oracle/make_golden.py [--out tests/golden] [--jobs 8]

Oracle definition: reference source at /root/reference executed under the
container's numpy / scipy (versions are recorded in every fixture).  The reference is only
*instrumented from outside* (its `solve_ivp` name is wrapped to record nfev/status); no reference
source is modified or copied.

Fixtures written (all small, committed):
  sim_raw.npz        Simulator6DOF raw mode (python-list IC/u => all-f64), incl. the
                     test_6DOF_simulator.py:3-7 known-answer input (dt=0.5).
  env_ka.npz         RNG-free env known-answer (ICRange=0) — three steps.
  constants.npz      derived constants of Rocket6DOF.__init__ for config.yaml.
  units.npz          action de-normalisation, reset quaternion rule, euler angles, quartic t_go.
  config1.npz        1 env, seed 42, 1000 random-action steps incl. the resets that occur.
  config2.npz        64 envs x 200 steps full trace  +  512 envs x 200 steps reward/done trace.
  policy_cl.npz      closed-loop best_model_2bo71j9m (numpy MLP), 30 episodes, + the MLP weights.
  velocity.npz       reward_shaping_type='velocity', 1 env x 400 steps.
  wrappers.npz       reward streams of the optional reward wrappers (wrappers.py:39-61, 128-155) replayed on
                     the first episodes of policy_cl: RewardAnnealing (make_annealed_env, main_6DOF.py:55-69)
                     and VerticalAttitudeReward on top of the plain and of the annealed env.
"""
import argparse
import copy
import io
import multiprocessing as mp
import os
import sys
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.ref_loader import REFERENCE_ROOT, load_reference  # noqa: E402

FLAG_NAMES = ["zero_height", "velocity_limit", "landing_radius", "attitude_limit", "omega_limit"]
TERM_NAMES = ["shaping", "thrust_penalty", "eta", "attitude_constraint", "goal_conditions",
              "final_position", "final_velocity"]


def _versions():
    import scipy
    return dict(numpy_version=np.__version__, scipy_version=scipy.__version__)


class Probe:
    """Wraps the `solve_ivp` name inside the reference's simulator module (instrumentation only)."""

    def __init__(self):
        import my_environment.utils.simulator as simmod
        self.simmod = simmod
        self.orig = simmod.solve_ivp
        self.nfev = -1
        self.status = 0
        simmod.solve_ivp = self._wrapped

    def _wrapped(self, *a, **k):
        sol = self.orig(*a, **k)
        self.nfev = sol.nfev
        self.status = sol.status
        return sol

    def close(self):
        self.simmod.solve_ivp = self.orig


def rollout(env_kwargs, seed, actions, policy=None, n_steps=None, max_episode_steps=1500):
    """Steps the reference env (auto-reset like a DummyVecEnv + TimeLimit) and records everything.

    actions: [T,3] float32 (open loop) or None with `policy` (callable obs13->action f32[3]).
    """
    Rocket6DOF, _, _, _ = load_reference()
    kw = copy.deepcopy(env_kwargs)
    kw["seed"] = seed
    env = Rocket6DOF(**kw)
    probe = Probe()
    T = n_steps if n_steps is not None else len(actions)
    rec = dict(
        ic=[], ic_step=[], state=np.zeros((T, 14)), obs=np.zeros((T, 14), np.float32),
        reward=np.zeros(T), terms=np.zeros((T, 7)), done=np.zeros(T, bool), oob=np.zeros(T, bool),
        status=np.zeros(T, np.int8), nfev=np.zeros(T, np.int16), flags=np.zeros((T, 5), bool),
        truncated=np.zeros(T, bool), action=np.zeros((T, 3), np.float32),
        u=np.zeros((T, 3), np.float32),
    )
    obs = env.reset()
    rec["ic"].append(env.initial_condition.copy())
    rec["ic_step"].append(0)
    ep_len = 0
    for k in range(T):
        a = np.float32(actions[k]) if policy is None else policy(obs[:13])
        rec["action"][k] = a
        obs, r, done, info = env.step(a)
        ep_len += 1
        rec["u"][k] = env.action
        rec["state"][k] = env.state
        rec["obs"][k] = obs
        rec["reward"][k] = r
        rec["terms"][k] = [float(v) for v in info["rewards_dict"].values()]
        rec["oob"][k] = info["bounds_violation"]
        rec["status"][k] = probe.status
        rec["nfev"][k] = probe.nfev
        rec["flags"][k] = [bool(v) for v in env._check_landing(env.state.astype(np.float32)).values()]
        trunc = (ep_len >= max_episode_steps) and not done
        rec["truncated"][k] = trunc
        rec["done"][k] = done
        if done or trunc:
            obs = env.reset()
            ep_len = 0
            rec["ic"].append(env.initial_condition.copy())
            rec["ic_step"].append(k + 1)
    probe.close()
    rec["ic"] = np.asarray(rec["ic"], np.float32)
    rec["ic_step"] = np.asarray(rec["ic_step"], np.int32)
    return rec


def _worker_cfg2(args):
    env_kwargs, seed, actions = args
    return rollout(env_kwargs, seed, actions)


def load_policy(name="best_model_2bo71j9m"):
    import torch
    with zipfile.ZipFile(os.path.join(REFERENCE_ROOT, name + ".zip")) as z:
        sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
    w = {k: v.numpy().astype(np.float32) for k, v in sd.items()}
    return dict(
        w0=w["mlp_extractor.shared_net.0.weight"], b0=w["mlp_extractor.shared_net.0.bias"],
        w1=w["mlp_extractor.shared_net.2.weight"], b1=w["mlp_extractor.shared_net.2.bias"],
        w2=w["action_net.weight"], b2=w["action_net.bias"],
    )


def mlp_forward(p, obs13):
    """Deterministic SB3 action: clip(action_net(tanh(L1(tanh(L0(obs))))), -1, 1), float32."""
    x = np.asarray(obs13, np.float32)
    h = np.tanh(p["w0"] @ x + p["b0"]).astype(np.float32)
    h = np.tanh(p["w1"] @ h + p["b1"]).astype(np.float32)
    a = (p["w2"] @ h + p["b2"]).astype(np.float32)
    return np.clip(a, -1.0, 1.0).astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden"))
    ap.add_argument("--jobs", type=int, default=os.cpu_count())
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    only = set(filter(None, args.only.split(",")))

    def want(name):
        return not only or name in only

    Rocket6DOF, Simulator6DOF, env_cfg, sb3_cfg = load_reference()
    ver = _versions()

    def save(name, **arrs):
        path = os.path.join(args.out, name + ".npz")
        np.savez_compressed(path, **arrs, **ver)
        print(f"wrote {path}  ({os.path.getsize(path) / 1024:.0f} KiB)")

    # ---------------------------------------------------------------- sim_raw
    if want("sim_raw"):
        ic0 = [100, 100, 100, 0, 3, 4, 1, 1, 0, 0, 0, 0, 0, 50e3]   # test_6DOF_simulator.py:3
        sim = Simulator6DOF(ic0)                                    # dt defaults to 0.5
        st, status = sim.step([0, 0, 1])
        ka_state, ka_status = np.array(st), status
        # a longer raw-mode run (dt = 0.1) with a varying f64 control list, until the ground event
        rng = np.random.default_rng(7)
        ic1 = [300.0, -50.0, 20.0, -40.0, 5.0, -2.0, 0.9, 0.1, -0.2, 0.3, 0.01, -0.02, 0.03, 45e3]
        sim = Simulator6DOF(ic1, 0.1)
        us, sts, stat = [], [], []
        for k in range(400):
            u = [float(rng.uniform(-0.3, 0.3)), float(rng.uniform(-0.3, 0.3)), float(rng.uniform(0, 981e3))]
            st, status = sim.step(u)
            us.append(u); sts.append(np.array(st)); stat.append(status)
            if status != 0:
                break
        save("sim_raw", ka_ic=np.array(ic0, float), ka_u=np.array([0., 0., 1.]), ka_dt=0.5,
             ka_state=ka_state, ka_status=ka_status,
             run_ic=np.array(ic1), run_dt=0.1, run_u=np.array(us), run_state=np.array(sts),
             run_status=np.array(stat, np.int8))

    # ---------------------------------------------------------------- env_ka
    if want("env_ka"):
        kw = copy.deepcopy(env_cfg)
        kw["ICRange"] = [0] * 14
        acts = np.float32([[0.3, -0.2, 0.5], [-1, 1, -1], [0, 0, 1]])
        rec = rollout(kw, 42, acts)
        save("env_ka", **rec)

    # ---------------------------------------------------------------- constants
    if want("constants"):
        env = Rocket6DOF(**copy.deepcopy(env_cfg))
        save("constants",
             ic_low=env.init_space.low, ic_high=env.init_space.high,
             max_gimbal=np.float64(env.max_gimbal), max_thrust=np.float64(env.max_thrust),
             state_normalizer=np.asarray(env.state_normalizer, np.float64),
             bounds_low=env.position_bounds_space.low, bounds_high=env.position_bounds_space.high,
             attitude_traj_limit=np.asarray(env.attitude_traj_limit, np.float64),
             landing_attitude_limit=np.asarray(env.landing_attitude_limit, np.float64),
             omega_lim=np.asarray(env.omega_lim, np.float64),
             max_episode_steps=np.int64(int(sb3_cfg["max_time"] / env_cfg["timestep"])))

    # ---------------------------------------------------------------- units
    if want("units"):
        from scipy.spatial.transform import Rotation as R
        rng = np.random.default_rng(11)
        env = Rocket6DOF(**copy.deepcopy(env_cfg))
        n = 4096
        a = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
        a[:8] = np.float32([[1, 1, 1], [-1, -1, -1], [0, 0, 0], [1, -1, 0], [0.5, 0.25, -0.75],
                            [-0.0, 0.0, -1.0], [1e-8, -1e-8, 1 - 1e-7], [0.999999, -0.999999, 0.333333]])
        u = np.stack([env._denormalize_action(x) for x in a])
        # per-step f32 constants of the simulator for those controls (rocket mass m0 random f32)
        m0 = rng.uniform(40500, 41500, n).astype(np.float32)
        tb = np.zeros((n, 3)); dm = np.zeros(n); J = np.zeros((n, 3)); Jinv = np.zeros((n, 3))
        for i in range(n):
            ic = np.zeros(14, np.float32); ic[6] = 1; ic[13] = m0[i]
            sim = Simulator6DOF(ic, 0.1)
            tb[i] = sim._get_thrust_body_frame(u[i])
            dm[i] = sim.RHS(0.0, ic.astype(np.float64), u[i])[13]
            J[i] = np.diag(sim.J); Jinv[i] = np.diag(sim.Jinv)
        # reset rule: sample -> f32 quaternion normalisation
        ics_raw = np.stack([env.init_space.sample() for _ in range(n)])
        ics = ics_raw.copy()
        for i in range(n):
            ics[i, 6:10] = ics[i, 6:10] / np.linalg.norm(ics[i, 6:10])
        # euler zyx of f32 quaternions (leading scalar), incl. near-gimbal-lock rows
        q = rng.normal(size=(n, 4)).astype(np.float32)
        q /= np.linalg.norm(q, axis=1, keepdims=True).astype(np.float32)
        th = rng.uniform(-np.pi, np.pi, 64)
        for i, t in enumerate(th):       # pitch = +-90 deg exactly / nearly
            s = 1 if i % 2 else -1
            qq = (R.from_euler("zyx", [t, s * (np.pi / 2 - (0 if i < 32 else 1e-9 * i)), 0.3 * t])).as_quat()
            q[i] = np.roll(qq, 1).astype(np.float32)
        eul = np.stack([R.from_quat(np.roll(x, -1)).as_euler("zyx") for x in q])
        # quartic t_go: np.roots selection of rocket_env.py:533-541 on (i) trajectory-like and
        # (ii) near-ground fast-dive inputs (3 positive roots)
        rv = np.zeros((n, 6), np.float32)
        rv[: n // 2, 0] = rng.uniform(0.5, 2100, n // 2); rv[: n // 2, 1] = rng.uniform(-2100, 2100, n // 2)
        rv[: n // 2, 2] = rng.uniform(-170, 170, n // 2); rv[: n // 2, 3] = rng.uniform(-250, 60, n // 2)
        rv[: n // 2, 4] = rng.uniform(-250, 250, n // 2); rv[: n // 2, 5] = rng.uniform(-40, 40, n // 2)
        k = n - n // 2
        rdir = rng.normal(size=(k, 3)); rdir /= np.linalg.norm(rdir, axis=1, keepdims=True)
        rmag = 10 ** rng.uniform(-1, 3, k)
        vmag = np.sqrt(6 * 9.81 * rmag) * rng.uniform(0.8, 4, k)
        vdir = -rdir + 0.15 * rng.normal(size=(k, 3)); vdir /= np.linalg.norm(vdir, axis=1, keepdims=True)
        rv[n // 2:, :3] = (rdir * rmag[:, None]).astype(np.float32)
        rv[n // 2:, 3:] = (vdir * vmag[:, None]).astype(np.float32)
        coef = np.zeros((n, 3)); tgo = np.full(n, np.nan); npos = np.zeros(n, np.int8)
        atarg = np.full((n, 3), np.nan)
        mass = rng.uniform(30000, 42000, n).astype(np.float32)
        for i in range(n):
            r, v = rv[i, :3], rv[i, 3:]
            c = [(-9.81) ** 2, 0, -4 * np.linalg.norm(v) ** 2, -24 * np.dot(r, v), -36 * np.linalg.norm(r) ** 2]
            coef[i] = [float(c[2]), float(c[3]), float(c[4])]
            sol = np.roots(c)
            pos = [s for s in sol if (s.imag == 0 and s.real > 0)]
            npos[i] = len(pos)
            if pos:
                tgo[i] = pos[0].real
                env.atarg_history = []
                atarg[i] = env._compute_atarg(r=np.array(r), v=np.array(v), mass=mass[i])
        save("units", act=a, act_u=u, m0=m0, tbody=tb, dm=dm, J=J, Jinv=Jinv,
             ic_raw=ics_raw, ic_norm=ics, quat=q, euler=eul,
             rv=rv, mass=mass, quartic_coef=coef, tgo=tgo, npos=npos, atarg=atarg)

    # ---------------------------------------------------------------- config1
    if want("config1"):
        acts = np.random.default_rng(0).uniform(-1, 1, (1000, 3)).astype(np.float32)
        rec = rollout(env_cfg, 42, acts)
        save("config1", **rec)

    # ---------------------------------------------------------------- config2
    if want("config2"):
        K, NF, NS = 200, 64, 512
        acts = np.random.default_rng(1).uniform(-1, 1, (K, NF + NS, 3)).astype(np.float32)
        jobs = [(env_cfg, 1000 + i, acts[:, i]) for i in range(NF + NS)]
        with mp.Pool(args.jobs) as pool:
            recs = pool.map(_worker_cfg2, jobs, chunksize=4)
        full, summ = recs[:NF], recs[NF:]

        def pack_ics(rs):
            m = max(len(r["ic"]) for r in rs)
            ic = np.zeros((len(rs), m, 14), np.float32); st = np.full((len(rs), m), -1, np.int32)
            for i, r in enumerate(rs):
                ic[i, : len(r["ic"])] = r["ic"]; st[i, : len(r["ic"])] = r["ic_step"]
            return ic, st
        icf, stf = pack_ics(full)
        ics, sts = pack_ics(summ)
        save("config2",
             actions=acts,
             full_ic=icf, full_ic_step=stf,
             full_state=np.stack([r["state"] for r in full], 1),
             full_obs=np.stack([r["obs"] for r in full], 1),
             full_reward=np.stack([r["reward"] for r in full], 1),
             full_terms=np.stack([r["terms"] for r in full], 1),
             full_done=np.stack([r["done"] for r in full], 1),
             full_oob=np.stack([r["oob"] for r in full], 1),
             full_status=np.stack([r["status"] for r in full], 1),
             full_nfev=np.stack([r["nfev"] for r in full], 1),
             full_flags=np.stack([r["flags"] for r in full], 1),
             summ_ic=ics, summ_ic_step=sts,
             summ_reward=np.stack([r["reward"] for r in summ], 1),
             summ_done=np.stack([r["done"] for r in summ], 1),
             summ_oob=np.stack([r["oob"] for r in summ], 1),
             summ_nfev=np.stack([r["nfev"] for r in summ], 1),
             summ_flags=np.stack([r["flags"] for r in summ], 1),
             summ_final_state=np.stack([r["state"][-1] for r in summ], 0))

    # ---------------------------------------------------------------- policy_cl
    if want("policy_cl"):
        p = load_policy()
        rec = rollout(env_cfg, 7, None, policy=lambda o: mlp_forward(p, o), n_steps=9000)
        # keep whole episodes only (montecarlo_script.py evaluates 30)
        ends = np.nonzero(rec["done"] | rec["truncated"])[0]
        n_ep = min(30, len(ends))
        T = int(ends[n_ep - 1]) + 1
        out = {k: (v[:T] if isinstance(v, np.ndarray) and v.shape[:1] == (9000,) else v) for k, v in rec.items()}
        out["ic"] = rec["ic"][:n_ep]; out["ic_step"] = rec["ic_step"][:n_ep]
        save("policy_cl", **out, **{"mlp_" + k: v for k, v in p.items()})

    # ---------------------------------------------------------------- wrappers
    if want("wrappers"):
        load_reference()
        from my_environment.wrappers.wrappers import RemoveMassFromObs, RewardAnnealing, VerticalAttitudeReward
        Rocket6DOF = load_reference()[0]
        g = np.load(os.path.join(args.out, "policy_cl.npz"))
        n_ep = 12
        T = int(g["ic_step"][n_ep])
        acts = g["action"][:T]

        def run(chain):
            kw = copy.deepcopy(env_cfg)
            kw["seed"] = 7                       # the seed policy_cl was recorded with: same ICs, same trajectories
            base = Rocket6DOF(**kw)
            env = chain(RemoveMassFromObs(base))
            rew = np.zeros(T)
            vert = np.zeros(T)
            thrust_pen = np.zeros(T)
            env.reset()
            for k in range(T):
                _, r, done, info = env.step(acts[k])
                rew[k] = r
                vert[k] = info["rewards_dict"].get("vertical_attitude_reward", 0)
                thrust_pen[k] = info["rewards_dict"]["thrust_penalty"]
                assert np.array_equal(base.state, g["state"][k]) and done == bool(g["done"][k])
                if done:
                    env.reset()
            return rew, vert, thrust_pen
        r_ann, _, tp_ann = run(lambda e: RewardAnnealing(e))
        r_va, v_a, _ = run(lambda e: VerticalAttitudeReward(RewardAnnealing(e)))
        r_vb, v_b, _ = run(lambda e: VerticalAttitudeReward(e))
        assert (v_b != 0).sum() > 0, "fixture does not exercise the vertical-attitude term"
        save("wrappers", n_steps=T, reward_annealed=r_ann, thrust_penalty_annealed=tp_ann,
             reward_vertical_annealed=r_va, vertical_term_annealed=v_a,
             reward_vertical_base=r_vb, vertical_term_base=v_b, xi=env_cfg["reward_coeff"].get("xi", 0.01), threshold_height=1e-3, weight=-0.5)

    # ---------------------------------------------------------------- velocity
    if want("velocity"):
        kw = copy.deepcopy(env_cfg)
        kw["reward_shaping_type"] = "velocity"
        acts = np.random.default_rng(5).uniform(-1, 1, (400, 3)).astype(np.float32)
        rec = rollout(kw, 43, acts)
        save("velocity", **rec)


if __name__ == "__main__":
    main()
