"""Parity of the path bench.py actually times ("What's weak" 1-2): the default dispatch of
`Rocket6DOFBatch` at >= 2^19 envs — multi-pass integrator (first pass | resume passes | tail) + post-step kernel on
two stream lanes — against the C oracle, and the multi-pass integrator on workloads where most envs need a second
and many a third RK attempt, so that every work-list path (long lists, several tiles per CTA) runs under the checker.

Reference behaviour: Rocket6DOF.step, /root/reference/my_environment/envs/rocket_env.py:201-231.
Bars as in test_gpu_parity.py: RK attempt counts / solver status / done / flags exact, state <= 1e-9, reward <= 1e-6.
"""
import numpy as np
import pytest

from parity_utils import RTOL_REWARD_TRAJ, RTOL_STATE, env_params, f32_ulp_diff, reward_err_traj, state_err

pytestmark = pytest.mark.gpu

LAND_BITS = [8, 16, 32, 64, 128]


def _fetch_idx(env, idx):
    """Everything a step produced, for the env indices `idx` (device tensor) only."""
    import torch
    torch.cuda.synchronize()
    fl = env.flags[idx].cpu().numpy()
    return dict(
        state=env.state[:, idx].t().cpu().numpy(), obs=env.obs[:, idx].t().cpu().numpy(),
        reward=env.reward[idx].cpu().numpy(), done=env.done[idx].cpu().numpy().astype(bool),
        event=(fl & 1) != 0, oob=(fl & 2) != 0, trunc=(fl & 4) != 0,
        flags=np.stack([(fl & b) != 0 for b in LAND_BITS], -1),
        natt=env.nattempts[idx].cpu().numpy().astype(np.int64), status=env.status[idx].cpu().numpy(),
        tstate=env.terminal_state[:, idx].t().cpu().numpy(), tobs=env.terminal_obs[:, idx].t().cpu().numpy(),
        v0=env.v0[idx].cpu().numpy(), terms=env.reward_terms[:, idx].t().cpu().numpy(),
    )


@pytest.mark.parametrize("lanes", [1, 2])
@pytest.mark.parametrize("dt", [1.0, 0.25])
def test_multipass_long_work_lists_vs_oracle(dt, lanes):
    """dt = 0.25 s: every env needs a second RK attempt and ~6 % a third; dt = 1 s: half of them need three and 7 %
    four to six — so the unfinished-env lists of the multi-pass integrator are as long as the range itself (the
    resume launches are sized for the random-action mix at dt = 0.1 s, where list 0 holds ~64 % and list 1 ~1 %) and
    each resume CTA walks more than one tile.  8192 envs x 40 (12 at dt = 1 s, where episodes are short) random-action
    steps against the C oracle."""
    import torch
    from oracle import c_oracle as co
    from rl_rocket_6dof_b200.batch import Rocket6DOFBatch
    ep = env_params(timestep=dt)
    n, K = 8192, (12 if dt == 1.0 else 40)
    env = Rocket6DOFBatch(n, params=ep, auto_reset=False, clip_reward=False, time_limit=False, debug_buffers=True,
                          split_step=True, multipass=True, lanes=lanes, seed=31)
    assert env.work is not None
    env.reset()
    torch.cuda.synchronize()
    ic = env.state.t().cpu().numpy()
    ob = co.OracleBatch(ep, n, nthreads=8)
    ob.set_state(ic, ic[:, 13].astype(np.float32), 0, v0=env.v0.cpu().numpy())
    rng = np.random.default_rng(12)
    alive = np.ones(n, bool)
    idx = torch.arange(n, device="cuda")
    hist = np.zeros(16, np.int64)
    for k in range(K):
        a = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
        env.step(torch.from_numpy(a).cuda())
        o = _fetch_idx(env, idx)
        r = ob.step(a)
        assert np.array_equal(2 + 6 * o["natt"][alive], r["nfev"][alive]), k
        assert np.array_equal(o["status"][alive], r["status"][alive]), k
        assert np.array_equal(o["done"][alive], r["done"][alive].astype(bool)), k
        assert np.array_equal(o["flags"][alive], r["flags"][alive].astype(bool)), k
        assert state_err(o["state"][alive], r["state"][alive], ep.state_normalizer).max() <= RTOL_STATE, k
        assert reward_err_traj(o["reward"][alive], r["reward"][alive]).max() <= RTOL_REWARD_TRAJ, k
        hist += np.bincount(o["natt"][alive], minlength=16)[:16]
        alive &= ~o["done"]
    frac = hist / hist.sum()
    print(f"dt={dt} lanes={lanes}: attempts histogram {np.round(frac[:6], 4)}, alive at the end {alive.sum()}")
    assert frac[2:].sum() > 0.75               # list 0 longer than 3/4 of the range
    if dt == 1.0:
        assert frac[3:].sum() > 0.25           # list 1 far longer than 1/16 of the range
        assert frac[4:].sum() > 0.02           # and a tail of 4-6 attempts
    assert alive.sum() > (0 if dt == 1.0 else n // 8)


def test_default_dispatch_1m_envs_sampled_vs_oracle():
    """BASELINE.json configs[2] exactly as bench.py builds it: 2^20 envs, default dispatch (kernel pair, multi-pass
    integrator from 2^19 envs, two stream lanes), auto-reset / ClipReward / TimeLimit on, after a pre-roll to the
    steady-state mix of episode phases.  8192 randomly chosen env indices are followed in lock-step by the C oracle
    for 64 steps (fresh initial conditions are read back from the device after each reset)."""
    import torch
    from oracle import c_oracle as co
    from rl_rocket_6dof_b200.batch import Rocket6DOFBatch
    ep = env_params()
    N, M, K = 1 << 20, 8192, 64
    env = Rocket6DOFBatch(N, params=ep, seed=42, debug_buffers=True, lanes=2)
    assert env.work is not None and env.scratch is not None and env.lanes == 2      # what bench.py times
    env.reset()
    env.rollout(150)                                                               # pre-roll (as bench.py does)
    rng = np.random.default_rng(2024)
    pick = np.sort(rng.choice(N, M, replace=False))
    pick[0], pick[-1] = 0, N - 1                                                   # both ends of the lane ranges
    half = N // 2
    pick[M // 2 - 1], pick[M // 2] = half - 1, half                                # the seam between the two lanes
    pick = np.unique(pick)
    M = len(pick)
    idx = torch.from_numpy(pick).cuda()
    torch.cuda.synchronize()
    ob = co.OracleBatch(ep, M, nthreads=8)
    y0 = env.state[:, idx].t().cpu().numpy()
    ob.set_state(y0, env.m0[idx].cpu().numpy(), env.step_count[idx].cpu().numpy(), v0=env.v0[idx].cpu().numpy())
    gen = torch.Generator(device="cuda"); gen.manual_seed(7)
    n_done = n_flip = 0
    worst = 0.0
    hist = np.zeros(8, np.int64)
    for k in range(K):
        a = (torch.rand(N, 3, device="cuda", generator=gen) * 2 - 1).contiguous()
        env.step(a)
        o = _fetch_idx(env, idx)
        r = ob.step(a[idx].cpu().numpy())
        done_ref = r["done"].astype(bool)
        assert not o["trunc"].any()
        assert np.array_equal(o["done"], done_ref), k
        assert np.array_equal(o["oob"], r["oob"].astype(bool)), k
        assert np.array_equal(o["status"], r["status"]), k
        assert np.array_equal(2 + 6 * o["natt"], r["nfev"]), k
        assert np.array_equal(o["flags"], r["flags"].astype(bool)), k
        hist += np.bincount(o["natt"], minlength=8)[:8]
        st, ob_ = o["state"].copy(), o["obs"].copy()
        d = np.nonzero(done_ref)[0]
        st[d], ob_[d] = o["tstate"][d], o["tobs"][d]
        se = state_err(st, r["state"], ep.state_normalizer).max()
        worst = max(worst, se)
        assert se <= RTOL_STATE, k
        ul = f32_ulp_diff(ob_[:, :13], r["obs"][:, :13])
        assert ul.max() <= 1, k
        n_flip += (ul >= 0.5).sum()
        ref_rew = np.clip(r["reward"], ep.clip_lo, ep.clip_hi)                     # ClipReward(-1, 100), main_6DOF.py:40-42
        assert reward_err_traj(o["reward"], ref_rew).max() <= RTOL_REWARD_TRAJ, k
        assert reward_err_traj(o["terms"], r["terms"]).max() <= RTOL_REWARD_TRAJ, k
        if len(d):
            n_done += len(d)
            new_ic = o["state"][d].astype(np.float32)
            assert np.array_equal(new_ic.astype(np.float64), o["state"][d])        # reset states are float32 values
            ob.set_state(o["state"][d], new_ic[:, 13], 0, v0=o["v0"][d], idx=d)
    frac = hist / hist.sum()
    print(f"2^20 default dispatch, {M} sampled envs x {K} steps: worst state err {worst:.2e}, episodes ended {n_done}, "
          f"obs ulp flips {n_flip}, attempts {np.round(frac[:5], 4)}")
    assert n_done > M // 8
    assert n_flip <= 400
    s = env.stats_dict()
    assert s["steps"] == N * (150 + K)
