"""Trajectory recorder (`SIM.states` / `SIM.actions` / `SIM.times`, simulator.py:30-35,100-102) on the GPU path."""
import numpy as np
import pytest

from parity_utils import RTOL_STATE, env_params, golden, state_err

pytestmark = pytest.mark.gpu


def _batch(n, **kw):
    from rl_rocket_6dof_b200.batch import Rocket6DOFBatch
    kw.setdefault("clip_reward", False)
    kw.setdefault("time_limit", False)
    return Rocket6DOFBatch(n, params=env_params(), **kw)


def test_record_first_config1_episode_against_reference_lists():
    """First episode of the config1 fixture (94 steps): the recorded lists are what the reference's simulator held
    at done — states bit-level parity bars, denormalised actions exact, times the rounded step clock."""
    import torch
    from rl_rocket_6dof_b200.recorder import TrajectoryRecorder
    g = golden("config1")
    L = int(np.nonzero(g["done"])[0][0]) + 1
    env = _batch(1, auto_reset=False)
    rec = TrajectoryRecorder(env, L)
    rec.reset()
    ic = torch.from_numpy(g["ic"][0:1])
    env.set_state(ic)
    rec.states[0].copy_(env.state)
    rec.initial[1].copy_(env.state)
    acts = torch.from_numpy(g["action"][:L]).cuda()
    for k in range(L):
        rec.step(acts[k:k + 1])
    e = rec.episode(0)
    assert e["states"].shape == (L + 1, 14) and e["actions"].shape == (L + 1, 3) and e["times"].shape == (L + 1,)
    assert np.array_equal(e["states"][0], g["ic"][0].astype(np.float64))
    ep = env_params()
    for k in range(L):
        assert state_err(e["states"][k + 1], g["state"][k], ep.state_normalizer) <= RTOL_STATE, k
    assert np.array_equal(e["actions"][0], np.zeros(3, np.float32))
    assert np.array_equal(e["actions"][1:], g["u"][:L])
    t_ref = [0.0]
    for _ in range(L):
        t_ref.append(round(t_ref[-1] + ep.timestep, 3))
    assert np.array_equal(e["times"], np.asarray(t_ref))
    df = rec.states_to_dataframe(0)
    assert list(df.columns)[-1] == "mass" and len(df) == L + 1
    assert rec.used_mass(0) == pytest.approx(float(g["ic"][0][-1]) - g["state"][L - 1][-1], rel=1e-9)
    with pytest.raises(RuntimeError):
        rec.step(acts[:1])


def test_record_auto_reset_episodes_are_self_consistent():
    """With auto-reset the record splits into episodes; each episode's rows replay exactly from its recorded initial
    condition and actions in a fresh batch, its last row is the terminal state, and its clock restarts."""
    import torch
    from rl_rocket_6dof_b200.recorder import record_rollout
    n, k = 64, 260
    env = _batch(n, auto_reset=True, seed=5)
    acts = torch.from_numpy(np.random.default_rng(3).uniform(-1, 1, (k, n, 3)).astype(np.float32)).cuda()
    rec = record_rollout(env, k, actions=acts)
    assert rec.length == k
    starts = rec.episode_start[1:k + 2].sum(0).cpu().numpy()
    assert (starts >= 2).all()                      # random-action episodes last 99-211 steps
    ep = env_params()
    checked = 0
    for i in (0, 17, 63):
        first = rec.episode(i, 0)
        second = rec.episode(i, 1)
        L0 = len(first["times"]) - 1
        assert first["times"][0] == 0.0 and second["times"][0] == 0.0
        assert second["times"][1] == pytest.approx(ep.timestep)
        # terminal row: on the ground (or out of bounds); next episode starts from a fresh initial condition
        assert not np.array_equal(first["states"][-1], second["states"][0])
        # replay episode 1 from its recorded initial condition
        one = _batch(1, auto_reset=False)
        one.reset()
        one.set_state(torch.from_numpy(second["states"][0:1]))
        L1 = len(second["times"]) - 1
        for j in range(L1):
            one.step(acts[L0 + j, i:i + 1].contiguous())
            torch.cuda.synchronize()
            got = one.state[:, 0].cpu().numpy()
            assert state_err(got, second["states"][j + 1], ep.state_normalizer) <= 1e-11, (i, j)
        checked += L1
    assert checked > 100


def test_record_rollout_with_policy():
    import torch
    from rl_rocket_6dof_b200 import policy as pol
    from rl_rocket_6dof_b200.recorder import record_rollout
    rng = np.random.default_rng(1)
    w = {"w0": rng.normal(0, 0.3, (128, 13)), "b0": rng.normal(0, 0.1, 128), "w1": rng.normal(0, 0.1, (64, 128)),
         "b1": rng.normal(0, 0.1, 64), "w2": rng.normal(0, 0.1, (3, 64)), "b2": rng.normal(0, 0.1, 3)}
    w = {k: v.astype(np.float32) for k, v in w.items()}
    env = _batch(32, auto_reset=True, seed=2)
    mlp = pol.to_device(w, env.device)
    rec = record_rollout(env, 20, mlp=mlp)
    a = rec.actions[1:21].cpu().numpy()
    assert np.isfinite(a).all() and (np.abs(a[..., :2]) <= np.deg2rad(20) + 1e-6).all() and (a[..., 2] >= 0).all()
    assert torch.isfinite(rec.states[:21]).all()
