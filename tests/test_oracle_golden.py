"""Pins the CPU oracle (oracle/r6_oracle.c) to the fixtures dumped from the unmodified reference
(oracle/make_golden.py).  CPU-only; this is the "oracle against golden vectors" gate of the tier."""
import numpy as np
import pytest

from oracle import c_oracle as co
from parity_utils import (RTOL_REWARD_TRAJ, RTOL_REWARD_UNIT, RTOL_STATE, env_params, f32_norm3,
                          f32_ulp_diff, golden, reward_err, reward_err_traj, state_err)


def replay(ep, rec, n_steps=None):
    """Steps one oracle env through a golden record, injecting the recorded initial conditions."""
    T = len(rec["action"]) if n_steps is None else n_steps
    ob = co.OracleBatch(ep, 1)
    ic_at = {int(s): rec["ic"][j] for j, s in enumerate(rec["ic_step"]) if s >= 0}
    out = {k: [] for k in ("state", "obs", "reward", "terms", "done", "oob", "status", "nfev", "flags")}
    for k in range(T):
        if k in ic_at:
            ic = ic_at[k]
            ob.set_state(ic.astype(np.float64), ic[13], 0, v0=f32_norm3(ic[3:6]))
        o = ob.step(rec["action"][k:k + 1])[0]
        for name in out:
            out[name].append(np.array(o[name]))
    return {k: np.array(v) for k, v in out.items()}


def check_record(ep, rec, n_steps=None, expect_obs_flips=8):
    T = len(rec["action"]) if n_steps is None else n_steps
    out = replay(ep, rec, T)
    assert np.array_equal(out["done"].astype(bool), rec["done"][:T])
    assert np.array_equal(out["oob"].astype(bool), rec["oob"][:T])
    assert np.array_equal(out["status"], rec["status"][:T])
    assert np.array_equal(out["nfev"], rec["nfev"][:T])
    assert np.array_equal(out["flags"].astype(bool), rec["flags"][:T])
    assert state_err(out["state"], rec["state"][:T], ep.state_normalizer).max() <= RTOL_STATE
    ulps = f32_ulp_diff(out["obs"], rec["obs"][:T])
    assert ulps.max() <= 1.0
    assert (ulps >= 0.5).sum() <= expect_obs_flips
    assert reward_err_traj(out["reward"], rec["reward"][:T]).max() <= RTOL_REWARD_TRAJ
    assert reward_err_traj(out["terms"], rec["terms"][:T]).max() <= RTOL_REWARD_TRAJ
    return out


def test_known_answer_test_6dof_simulator():
    """/root/reference/test_6DOF_simulator.py:3-7 (dt = 0.5, python-list inputs => raw f64 mode)."""
    g = golden("sim_raw")
    y, st, nfev = co.sim_step_raw(g["ka_ic"], g["ka_u"], g["ka_ic"][13], 0.0, float(g["ka_dt"]))
    assert st == int(g["ka_status"]) == 0
    assert np.max(np.abs(y - g["ka_state"]) / np.maximum(np.abs(g["ka_state"]), 1e-9)) <= 1e-12
    # SURVEY §D.1 literal
    assert abs(y[0] - 98.773873474135954) < 1e-11 and abs(y[13] - 49999.999858421106) < 1e-8


def test_raw_simulator_run_to_ground_event():
    g = golden("sim_raw")
    y, t = g["run_ic"].copy(), 0.0
    for k in range(len(g["run_u"])):
        y, st, _ = co.sim_step_raw(y, g["run_u"][k], g["run_ic"][13], t, float(g["run_dt"]))
        t = round(t + float(g["run_dt"]), 3)
        assert st == g["run_status"][k]
        assert np.max(np.abs(y - g["run_state"][k]) / np.maximum(np.abs(g["run_state"][k]), 1e-3)) <= 1e-11
    assert st == 1   # the run ends on the terminal height event


def test_env_known_answer_rng_free():
    """config.yaml with ICRange = 0, three steps."""
    ep = env_params(ICRange=[0] * 14)
    rec = golden("env_ka")
    out = check_record(ep, rec, expect_obs_flips=0)
    assert abs(out["reward"][0] - 0.05063647241362953) < 1e-15
    assert reward_err(out["reward"], rec["reward"]).max() <= RTOL_REWARD_UNIT


def test_units_action_and_step_constants():
    u = golden("units")
    p = co.make_params(env_params())
    for i in range(len(u["act"])):
        uu = co.denormalize_action(p, u["act"][i])
        assert np.array_equal(uu, u["act_u"][i]), i
    for i in range(0, len(u["act"]), 4):
        c = co.step_consts(u["m0"][i], u["act_u"][i])
        assert np.array_equal(c[0:2], u["J"][i][[0, 1]]) and np.array_equal(c[2:4], u["Jinv"][i][[0, 1]])
        assert np.array_equal(c[4:7], u["tbody"][i]), i
        assert c[7] == u["dm"][i]


def test_units_reset_quaternion_rule():
    u = golden("units")
    ob = co.OracleBatch(env_params(), len(u["ic_raw"]))
    obs, ic = ob.reset_from_samples(u["ic_raw"])
    assert np.array_equal(ic, u["ic_norm"])


def test_units_euler_zyx():
    u = golden("units")
    for i in range(len(u["quat"])):
        e = co.euler_zyx(u["quat"][i].astype(np.float64))
        d = np.abs(e - u["euler"][i])
        d = np.minimum(d, 2 * np.pi - d)
        assert d.max() < 5e-9 if i < 64 else d.max() < 1e-13, (i, e, u["euler"][i])


def test_units_tgo_quartic_vs_np_roots():
    u = golden("units")
    n3 = 0
    for i in range(len(u["tgo"])):
        t, npos = co.tgo(*u["quartic_coef"][i])
        assert npos == u["npos"][i]
        if npos:
            assert abs(t - u["tgo"][i]) <= 1e-13 * u["tgo"][i], (i, t, u["tgo"][i])
            n3 += npos == 3
    assert n3 > 100     # the three-positive-root branch (largest root) is exercised


def test_config1_single_env_1000_random_steps():
    check_record(env_params(), golden("config1"))


def test_velocity_shaping():
    check_record(env_params(reward_shaping_type="velocity"), golden("velocity"))


def test_policy_closed_loop_replay():
    rec = golden("policy_cl")
    check_record(env_params(), rec, expect_obs_flips=40)


def test_config2_batched_subset():
    g = golden("config2")
    ep = env_params()
    K, NF = g["full_state"].shape[:2]
    ob = co.OracleBatch(ep, NF, nthreads=4)
    n_flip = 0
    for k in range(K):
        for i in range(NF):
            j = np.nonzero(g["full_ic_step"][i] == k)[0]
            if len(j):
                ic = g["full_ic"][i, j[0]]
                ob.set_state(ic.astype(np.float64), ic[13], 0, v0=f32_norm3(ic[3:6]), idx=i)
        o = ob.step(g["actions"][k, :NF])
        assert np.array_equal(o["done"].astype(bool), g["full_done"][k])
        assert np.array_equal(o["oob"].astype(bool), g["full_oob"][k])
        assert np.array_equal(o["status"], g["full_status"][k])
        assert np.array_equal(o["nfev"], g["full_nfev"][k])
        assert np.array_equal(o["flags"].astype(bool), g["full_flags"][k])
        assert state_err(o["state"], g["full_state"][k], ep.state_normalizer).max() <= RTOL_STATE
        ul = f32_ulp_diff(o["obs"], g["full_obs"][k])
        assert ul.max() <= 1
        n_flip += (ul >= 0.5).sum()
        assert reward_err_traj(o["reward"], g["full_reward"][k]).max() <= RTOL_REWARD_TRAJ
    assert n_flip <= 64
    # reward / done trace of the 512 summary envs
    NS = g["summ_reward"].shape[1]
    ob = co.OracleBatch(ep, NS, nthreads=4)
    for k in range(K):
        hit = np.argwhere(g["summ_ic_step"] == k)
        for i, j in hit:
            ic = g["summ_ic"][i, j]
            ob.set_state(ic.astype(np.float64), ic[13], 0, v0=f32_norm3(ic[3:6]), idx=i)
        o = ob.step(g["actions"][k, NF:NF + NS])
        assert np.array_equal(o["done"].astype(bool), g["summ_done"][k])
        assert np.array_equal(o["nfev"], g["summ_nfev"][k])
        assert np.array_equal(o["flags"].astype(bool), g["summ_flags"][k])
        assert reward_err_traj(o["reward"], g["summ_reward"][k]).max() <= RTOL_REWARD_TRAJ
    assert state_err(o["state"], g["summ_final_state"], ep.state_normalizer).max() <= RTOL_STATE
