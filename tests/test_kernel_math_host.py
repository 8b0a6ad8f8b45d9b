"""CPU check of the CUDA kernels' per-environment code (tests/hostsim = g++ build of
rl_rocket_6dof_b200/csrc/r6_core.cuh) against the reference fixtures.  This is a development gate
that needs no GPU; the parity tests proper (-m gpu) run the same cases through libr6dof.so."""
import ctypes as C

import numpy as np

import hostsim
from parity_utils import (RTOL_REWARD_TRAJ, RTOL_STATE, env_params, f32_norm3, f32_ulp_diff, golden,
                          reward_err_traj, state_err)

F_EVENT, F_OOB, F_TRUNC = 1, 2, 4
LAND_BITS = [8, 16, 32, 64, 128]


def unpack(o):
    fl = o["flags"]
    return dict(done=(fl & (F_EVENT | F_OOB)) != 0, oob=(fl & F_OOB) != 0,
                flags=np.stack([(fl & b) != 0 for b in LAND_BITS], -1), nfev=2 + 6 * o["natt"])


def check_record(ep, rec, flips=8):
    T = len(rec["action"])
    hb = hostsim.HostSimBatch(ep, 1, auto_reset=False, clip_reward=False, time_limit=False)
    ic_at = {int(s): rec["ic"][j] for j, s in enumerate(rec["ic_step"]) if s >= 0}
    nflip = 0
    for k in range(T):
        if k in ic_at:
            ic = ic_at[k]
            hb.set_state(ic.astype(np.float64), ic[13], 0, v0=f32_norm3(ic[3:6]))
        o = hb.step(rec["action"][k:k + 1])
        u = unpack(o)
        assert u["done"][0] == rec["done"][k], k
        assert u["oob"][0] == rec["oob"][k], k
        assert o["status"][0] == rec["status"][k], k
        assert u["nfev"][0] == rec["nfev"][k], k
        assert np.array_equal(u["flags"][0], rec["flags"][k]), k
        assert o["tgo_missing"][0] == 0
        assert state_err(o["state"][0], rec["state"][k], ep.state_normalizer) <= RTOL_STATE, k
        ul = f32_ulp_diff(o["obs"][0], rec["obs"][k])
        assert ul.max() <= 1.0
        nflip += (ul >= 0.5).sum()
        assert reward_err_traj(o["reward"][0], rec["reward"][k]) <= RTOL_REWARD_TRAJ, k
        assert reward_err_traj(o["terms"][0], rec["terms"][k]).max() <= RTOL_REWARD_TRAJ, k
    assert nflip <= flips


def test_host_env_known_answer():
    check_record(env_params(ICRange=[0] * 14), golden("env_ka"), flips=0)


def test_host_config1():
    check_record(env_params(), golden("config1"))


def test_host_velocity_shaping():
    check_record(env_params(reward_shaping_type="velocity"), golden("velocity"))


def test_host_policy_closed_loop_replay():
    check_record(env_params(), golden("policy_cl"), flips=40)


def test_host_config2_full_trace():
    g = golden("config2")
    ep = env_params()
    K, NF = g["full_state"].shape[:2]
    hb = hostsim.HostSimBatch(ep, NF, auto_reset=False, clip_reward=False, time_limit=False)
    for k in range(K):
        for i, j in np.argwhere(g["full_ic_step"] == k):
            ic = g["full_ic"][i, j]
            hb.set_state(ic.astype(np.float64), ic[13], 0, v0=f32_norm3(ic[3:6]), idx=i)
        o = hb.step(g["actions"][k, :NF])
        u = unpack(o)
        assert np.array_equal(u["done"], g["full_done"][k])
        assert np.array_equal(u["oob"], g["full_oob"][k])
        assert np.array_equal(o["status"], g["full_status"][k])
        assert np.array_equal(u["nfev"], g["full_nfev"][k])
        assert np.array_equal(u["flags"], g["full_flags"][k])
        assert state_err(o["state"], g["full_state"][k], ep.state_normalizer).max() <= RTOL_STATE
        assert f32_ulp_diff(o["obs"], g["full_obs"][k]).max() <= 1
        assert reward_err_traj(o["reward"], g["full_reward"][k]).max() <= RTOL_REWARD_TRAJ


def test_host_raw_simulator():
    g = golden("sim_raw")
    L = hostsim.lib()
    y = g["ka_ic"].copy()
    natt = C.c_int(0)
    st = L.hs_sim_step_raw(y.ctypes.data, g["ka_u"].ctypes.data, float(g["ka_ic"][13]), 0.0, 0.5, C.byref(natt))
    assert st == 0
    assert np.max(np.abs(y - g["ka_state"]) / np.maximum(np.abs(g["ka_state"]), 1e-9)) <= 1e-11
    y, t = g["run_ic"].copy(), 0.0
    for k in range(len(g["run_u"])):
        u = np.ascontiguousarray(g["run_u"][k])
        st = L.hs_sim_step_raw(y.ctypes.data, u.ctypes.data, float(g["run_ic"][13]), t, 0.1, C.byref(natt))
        t = round(t + 0.1, 3)
        assert st == g["run_status"][k]
        assert np.max(np.abs(y - g["run_state"][k]) / np.maximum(np.abs(g["run_state"][k]), 1e-3)) <= 1e-10


def test_host_integrator_cut_into_passes_is_bit_identical():
    """integrate<kPass = 1 / 2> (one RK attempt per pass, only y + PassCtx carried, as the multi-pass kernels do) ==
    the single-call integrator bit for bit, on the config1 trajectory (1000 recorded states / controls, including the
    touchdown steps with the event path and the rare 3-attempt steps)."""
    g = golden("config1")
    L = hostsim.lib()
    m0 = {int(s): float(ic[13]) for s, ic in zip(g["ic_step"], g["ic"]) if s >= 0}
    cur_m0, npass = None, {}
    n3 = 0
    for k in range(len(g["action"]) - 1):
        if k in m0:
            cur_m0 = m0[k]
        if g["done"][k]:
            continue                                   # state[k] is terminal: the next record starts from a new IC
        y0 = np.ascontiguousarray(g["state"][k], np.float64)
        u = np.ascontiguousarray(g["u"][k + 1], np.float64)
        t = round(0.1 * (k % 50), 3)
        ya, yb = y0.copy(), y0.copy()
        na, nb, ps = C.c_int(0), C.c_int(0), C.c_int(0)
        sa = L.hs_sim_step_raw(ya.ctypes.data, u.ctypes.data, cur_m0, t, 0.1, C.byref(na))
        sb = L.hs_sim_step_raw_passes(yb.ctypes.data, u.ctypes.data, cur_m0, t, 0.1, C.byref(nb), C.byref(ps))
        assert sa == sb and na.value == nb.value and ps.value == nb.value, k
        assert np.array_equal(ya, yb), k
        npass[ps.value] = npass.get(ps.value, 0) + 1
        n3 += nb.value >= 3
    assert npass.get(1, 0) > 50 and npass.get(2, 0) > 50, npass
    # a longer step makes every call multi-pass, with rejected attempts and ground contact among them
    rng = np.random.default_rng(0)
    many = 0
    for k in rng.choice(len(g["action"]) - 1, 200, replace=False):
        y0 = np.ascontiguousarray(g["state"][k], np.float64)
        u = np.ascontiguousarray(g["u"][k], np.float64)
        ya, yb = y0.copy(), y0.copy()
        na, nb, ps = C.c_int(0), C.c_int(0), C.c_int(0)
        sa = L.hs_sim_step_raw(ya.ctypes.data, u.ctypes.data, 41000.0, 0.0, 2.0, C.byref(na))
        sb = L.hs_sim_step_raw_passes(yb.ctypes.data, u.ctypes.data, 41000.0, 0.0, 2.0, C.byref(nb), C.byref(ps))
        assert sa == sb and na.value == nb.value == ps.value and np.array_equal(ya, yb), k
        many += nb.value >= 3
    assert many > 20


def test_host_tgo_vs_np_roots():
    u = golden("units")
    L = hostsim.lib()
    for i in range(len(u["tgo"])):
        t = L.hs_tgo((-9.81) ** 2, *[float(x) for x in u["quartic_coef"][i]])
        if u["npos"][i]:
            assert abs(t - u["tgo"][i]) <= 1e-13 * u["tgo"][i], (i, t, u["tgo"][i], u["npos"][i])
        else:
            assert not (t > 0)
    # fuzz against np.roots on fresh coefficient sets, incl. the three-positive-root region
    rng = np.random.default_rng(3)
    n3 = 0
    for _ in range(3000):
        r = 10 ** rng.uniform(-1, 3.3)
        v = np.sqrt(6 * 9.81 * r) * rng.uniform(0.3, 5)
        cosang = rng.uniform(-1, 1) if rng.random() < 0.5 else -rng.uniform(0.9, 1)
        c = [(-9.81) ** 2, 0.0, -4 * v * v, -24 * r * v * cosang, -36 * r * r]
        roots = np.roots(c)
        pos = [z.real for z in roots if z.imag == 0 and z.real > 0]
        if not pos:
            continue
        n3 += len(pos) == 3
        t = L.hs_tgo(*[c[0], c[2], c[3], c[4]])
        assert abs(t - pos[0]) <= 2e-13 * pos[0], (c, t, pos)
        assert pos[0] == max(pos)
        # warm starts: near the root, on either side, near a SMALLER positive root, useless — always the largest root
        for g in (pos[0] * 1.01, pos[0] * 0.99, pos[0] * 1.5, pos[0] * 0.6, min(pos) * 1.001, min(pos) * 0.999,
                  pos[0] * 1e-4, pos[0] * 1e4):
            tw = L.hs_tgo_warm(c[0], c[2], c[3], c[4], g)
            assert abs(tw - pos[0]) <= 2e-13 * pos[0], (c, g, tw, pos)
    assert n3 > 20


def test_host_euler_limit_tests_vs_angles():
    u = golden("units")
    L = hostsim.lib()
    rng = np.random.default_rng(5)
    for trial, (viol, land) in enumerate([([85, 85, 360], [10, 10, 360]), ([30, 20, 100], [60, 40, 90]),
                                          ([170, 89, 179], [5, 90, 180]), ([180, 90, 180], [0, 0, 0])]):
        viol = np.deg2rad(viol).astype(np.float64)
        land = np.deg2rad(land).astype(np.float64)
        for i in range(len(u["quat"])):
            e = u["euler"][i]
            if np.min(np.abs(np.abs(e)[:, None] - np.stack([viol, land], 1))) < 1e-9:
                continue    # on a threshold
            q = u["quat"][i].astype(np.float64)
            v, l = C.c_int(0), C.c_int(0)
            L.hs_euler_tests(viol.ctypes.data, land.ctypes.data, q.ctypes.data, C.byref(v), C.byref(l))
            assert bool(v.value) == bool(np.any(np.abs(e) > viol)), (trial, i, e)
            assert bool(l.value) == bool(np.any(np.abs(e) < land)), (trial, i, e)
