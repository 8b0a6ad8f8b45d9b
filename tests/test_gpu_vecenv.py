"""The reference-facing Python layer on the GPU: SB3 VecEnv protocol (auto-reset, info keys), the
pinned-host fast path with and without zero-copy, and the single-env gym facade."""
import numpy as np
import pytest

from parity_utils import env_params, f32_ulp_diff, golden, reward_err

pytestmark = pytest.mark.gpu


def _actions(n, k, seed=0):
    return np.random.default_rng(seed).uniform(-1, 1, (k, n, 3)).astype(np.float32)


def test_zero_copy_step_host_equals_staged_copies():
    """Kernel writing straight into mapped pinned memory == kernels + explicit D2H copies.  The zero-copy path runs
    the fused step kernel and the staged path the integrator | post-step pair, so the two agree to float64
    round-off (different FMA contraction), with identical episode boundaries."""
    import torch
    from rl_rocket_6dof_b200 import make_vec_env
    n, k = 4096, 220
    a = make_vec_env(n, device="cuda:0", seed=5, zero_copy=True)
    b = make_vec_env(n, device="cuda:0", seed=5, zero_copy=False)
    oa, ob = a.reset_host().copy(), b.reset_host().copy()
    assert oa.shape == (n, 13) and oa.dtype == np.float32 and np.array_equal(oa, ob)
    acts = _actions(n, k)
    pinned = torch.from_numpy(acts).pin_memory()
    n_done = 0
    for j in range(k):
        ra = a.step_host(pinned[j])              # pinned input: no staging copy at all
        rb = b.step_host(acts[j])                # numpy input: staged through the pinned buffer
        assert np.array_equal(ra[2], rb[2]), j                              # dones
        assert f32_ulp_diff(ra[0], rb[0]).max() <= 1.0, j                   # float32 observations
        assert np.abs(ra[1] - rb[1]).max() <= 1e-5, j                       # float32 rewards
        n_done += int(ra[2].sum())
    assert n_done > n // 2                       # episodes ended and were auto-reset on both paths
    norm = torch.as_tensor(a.batch.params.state_normalizer, device="cuda")[:, None]
    assert float(((a.batch.state - b.batch.state).abs() / norm).max()) <= 1e-11
    assert torch.equal(a.batch.episode_id, b.batch.episode_id) and torch.equal(a.batch.step_count, b.batch.step_count)
    assert a.h2d_bytes_per_step == n * 12 and a.d2h_bytes_per_step == n * (13 * 4 + 4 + 2)


def test_vecenv_protocol_and_infos():
    from rl_rocket_6dof_b200 import make_vec_env
    from rl_rocket_6dof_b200.vec_env import MonitorTag
    n, k = 64, 260
    env = make_vec_env(n, device="cuda:0", seed=9)
    assert env.num_envs == n and env.observation_space.shape == (13,) and env.action_space.shape == (3,)
    assert env.env_is_wrapped(MonitorTag) == [True] * n
    assert env.get_attr("max_thrust", [0, 1]) == [981e3, 981e3]
    with pytest.raises(AttributeError):
        env.get_attr("nope")
    obs = env.reset()
    assert obs.shape == (n, 13) and obs.dtype == np.float32 and obs.flags["C_CONTIGUOUS"]
    norm = env.batch.params.state_normalizer
    ret = np.zeros(n)
    length = np.zeros(n, np.int64)
    finished = 0
    acts = _actions(n, k, seed=3)
    for j in range(k):
        env.step_async(acts[j])
        obs, rews, dones, infos = env.step_wait()
        assert obs.shape == (n, 13) and rews.dtype == np.float32 and dones.dtype == np.bool_ and len(infos) == n
        assert np.all(rews >= -1.0) and np.all(rews <= 100.0)          # ClipReward(-1, 100)
        ret += rews
        length += 1
        for i in np.nonzero(dones)[0]:
            info = infos[i]
            finished += 1
            assert set(info) >= {"terminal_observation", "episode", "TimeLimit.truncated", "state_history"}
            assert info["episode"]["l"] == length[i]
            assert abs(info["episode"]["r"] - ret[i]) <= 1e-3 * max(1.0, abs(ret[i]))
            assert info["episode"]["r"] == round(info["episode"]["r"], 6)      # Monitor: float64 sum rounded to 6 decimals
            ts = info["state_history"][-1]
            assert ts.shape == (14,)
            tob = (ts / norm).astype(np.float32)[:13]
            assert f32_ulp_diff(info["terminal_observation"], tob).max() <= 1.0
            assert info["is_done"] != info["TimeLimit.truncated"]
            assert info["bounds_violation"] or ts[0] <= 1e-2 or info["TimeLimit.truncated"]
            # the returned obs is already the first observation of the next episode
            assert not np.array_equal(obs[i], info["terminal_observation"])
            ret[i] = 0
            length[i] = 0
        for i in np.nonzero(~dones)[0][:4]:
            assert infos[i] == {}
            infos[i]["scribble"] = j                      # a wrapper writing into an info dict must not leak into later steps
    assert finished >= n
    env.close()


def test_gym_facade_known_answer():
    """Rocket6DOF(**env_config) with ICRange = 0 reproduces the RNG-free three-step fixture."""
    from rl_rocket_6dof_b200 import Rocket6DOF, load_config
    _, cfg = load_config()
    cfg = {**cfg, "ICRange": [0] * 14}
    rec = golden("env_ka")
    env = Rocket6DOF(**cfg, device="cuda:0")
    obs0 = env.reset()
    assert obs0.shape == (14,) and obs0.dtype == np.float32
    assert np.array_equal(env.get_state(), rec["ic"][0])
    for k in range(3):
        obs, reward, done, info = env.step(rec["action"][k])
        assert f32_ulp_diff(obs, rec["obs"][k]).max() < 0.5
        assert reward_err(reward, rec["reward"][k]) <= 1e-9
        assert done == bool(rec["done"][k]) and info["bounds_violation"] == bool(rec["oob"][k])
        assert len(info["state_history"]) == k + 2 and len(info["rewards_dict"]) == 7
    assert abs(env.used_mass() - (rec["ic"][0][13] - rec["state"][2][13])) <= 1e-6
    with pytest.raises(KeyError):
        Rocket6DOF(IC=cfg["IC"], ICRange=cfg["ICRange"], device="cuda:0")     # no landing_params -> 'waypoint'


@pytest.mark.parametrize("zero_copy", [True, False])
def test_step_async_returns_before_the_step_is_done(zero_copy):
    """step_async only enqueues (SubprocVecEnv contract); step_wait delivers what step() would have; the caller may
    overwrite its action array in between."""
    from rl_rocket_6dof_b200 import make_vec_env
    n, k = 2048, 40
    a = _actions(n, k, seed=5)
    e1 = make_vec_env(n, seed=9, zero_copy=zero_copy)
    e2 = make_vec_env(n, seed=9, zero_copy=zero_copy)
    o1, o2 = e1.reset(), e2.reset()
    assert np.array_equal(o1, o2)
    with pytest.raises(RuntimeError):
        e1.step_wait()
    for j in range(k):
        buf = a[j].copy()
        e1.step_async(buf)
        buf[:] = 7.0                                   # the env must not still be reading the caller's array
        r1 = e1.step_wait()
        r2 = e2.step(a[j])
        assert np.array_equal(r1[0], r2[0]) and np.array_equal(r1[1], r2[1]) and np.array_equal(r1[2], r2[2])
        assert [sorted(d) for d in r1[3]] == [sorted(d) for d in r2[3]]
