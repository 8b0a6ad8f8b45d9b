"""Optional reward wrappers of the reference (RewardAnnealing, VerticalAttitudeReward;
my_environment/wrappers/wrappers.py:39-61, 128-155) fused as reward modes, against reward streams the
unmodified wrappers produced on the first 12 closed-loop episodes of policy_cl (tests/golden/wrappers.npz)."""
import numpy as np
import pytest

from parity_utils import RTOL_REWARD_TRAJ, env_params, f32_norm3, golden, reward_err_traj

MODES = [
    (dict(reward_annealing=True), "reward_annealed"),
    (dict(reward_annealing=True, vertical_attitude_reward=(1e-3, -0.5)), "reward_vertical_annealed"),
    (dict(vertical_attitude_reward=(1e-3, -0.5)), "reward_vertical_base"),
]


def _ic_at(rec, T):
    return {int(s): rec["ic"][j] for j, s in enumerate(rec["ic_step"]) if 0 <= s < T}


@pytest.mark.parametrize("kw,key", MODES)
def test_host_build_reward_modes(kw, key):
    import hostsim
    rec, w = golden("policy_cl"), golden("wrappers")
    T = int(w["n_steps"])
    ep = env_params()
    hb = hostsim.HostSimBatch(ep, 1, auto_reset=False, clip_reward=False, time_limit=False, **kw)
    ic_at = _ic_at(rec, T)
    hits = 0
    for k in range(T):
        if k in ic_at:
            ic = ic_at[k]
            hb.set_state(ic.astype(np.float64), ic[13], 0, v0=f32_norm3(ic[3:6]))
        o = hb.step(rec["action"][k:k + 1])
        assert reward_err_traj(o["reward"][0], w[key][k]) <= RTOL_REWARD_TRAJ, (k, o["reward"][0], w[key][k])
        if "annealing" in str(kw):
            assert o["terms"][0][1] == w["thrust_penalty_annealed"][k]          # float32 product, bit-exact
        hits += w[key][k] != rec["reward"][k]
    assert hits > 0


def test_struct_defaults_and_xi_from_config():
    ep = env_params()
    p = ep.to_struct()
    assert p.reward_mode == 0 and p.xi == np.float32(ep.reward_coeff.get("xi", 0.01))   # config.yaml: xi = 0.004
    cfg_ep = env_params(reward_coeff={**ep.reward_coeff, "xi": 0.25})
    p = cfg_ep.to_struct(reward_annealing=True, vertical_attitude_reward=(2e-3, -1.0))
    assert p.reward_mode == 3 and p.xi == np.float32(0.25) and p.va_threshold == 2e-3 and p.va_weight == -1.0


@pytest.mark.gpu
@pytest.mark.parametrize("kw,key", MODES)
def test_gpu_reward_modes(kw, key):
    import torch
    from rl_rocket_6dof_b200.batch import Rocket6DOFBatch
    rec, w = golden("policy_cl"), golden("wrappers")
    T = int(w["n_steps"])
    env = Rocket6DOFBatch(1, params=env_params(), device="cuda:0", auto_reset=False, clip_reward=False,
                          time_limit=False, **kw)
    ic_at = _ic_at(rec, T)
    acts = torch.from_numpy(rec["action"][:T]).cuda()
    rewards = torch.zeros(T, dtype=torch.float64, device="cuda")
    for k in range(T):
        if k in ic_at:
            env.set_state(torch.from_numpy(ic_at[k][None]))
        env.step(acts[k:k + 1])
        rewards[k] = env.reward[0]
    r = rewards.cpu().numpy()
    assert reward_err_traj(r, w[key][:T]).max() <= RTOL_REWARD_TRAJ


@pytest.mark.gpu
def test_annealed_vec_env_factory():
    from rl_rocket_6dof_b200 import make_annealed_vec_env
    env = make_annealed_vec_env(32, device="cuda:0", seed=1)
    env.reset()
    _, rews, _, _ = env.step(np.zeros((32, 3), np.float32))
    # mid-flight: only the thrust penalty -xi*(0+1) is non-zero, and there is no ClipReward(-1, 100)
    assert np.allclose(rews, -env.batch.params.reward_coeff["xi"], atol=1e-7) and env.reward_range == (-np.inf, np.inf)
