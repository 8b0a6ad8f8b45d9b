"""GAE scan kernel (r6_gae) against the NumPy restatement of SB3's recursion, plus closed forms."""
import numpy as np
import pytest

from oracle import gae_oracle


def _case(T, n, seed, p_done=0.05):
    rng = np.random.default_rng(seed)
    rew = rng.normal(0, 1, (T, n)).astype(np.float32)
    val = rng.normal(0, 3, (T, n)).astype(np.float32)
    done = rng.random((T, n)) < p_done
    last = rng.normal(0, 3, n).astype(np.float32)
    return rew, val, done, last


def test_oracle_closed_forms():
    # gamma = lambda = 1, no episode ends: returns are reward-to-go plus the bootstrap value
    rew, val, done, last = _case(7, 5, 0, p_done=0.0)
    adv, ret = gae_oracle.compute_returns_and_advantage(rew, val, done, last, 1.0, 1.0)
    rtg = np.cumsum(rew[::-1].astype(np.float64), 0)[::-1] + last
    assert np.allclose(ret, rtg, atol=1e-4)
    # lambda = 0: one-step TD error
    adv0, _ = gae_oracle.compute_returns_and_advantage(rew, val, done, last, 0.9, 0.0)
    nxt = np.concatenate([val[1:], last[None]], 0)
    assert np.allclose(adv0, rew + 0.9 * nxt - val, atol=1e-5)
    # an episode end cuts both the bootstrap and the recursion
    done[3] = True
    adv1, _ = gae_oracle.compute_returns_and_advantage(rew, val, done, last, 0.9, 0.8)
    assert np.allclose(adv1[3], rew[3] - val[3], atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("T,n", [(1, 1), (5, 33), (128, 4097), (64, 1 << 16)])
def test_gae_kernel_bit_exact(T, n):
    import torch
    from rl_rocket_6dof_b200.gae import compute_gae
    rew, val, done, last = _case(T, n, T * 1000 + n)
    a_ref, r_ref = gae_oracle.compute_returns_and_advantage(rew, val, done, last, 0.99, 0.95)
    a, r = compute_gae(torch.from_numpy(rew).cuda(), torch.from_numpy(val).cuda(), torch.from_numpy(done).cuda(),
                       torch.from_numpy(last).cuda(), 0.99, 0.95)
    assert np.array_equal(a.cpu().numpy(), a_ref) and np.array_equal(r.cpu().numpy(), r_ref)


@pytest.mark.gpu
def test_gae_on_recorded_rollout_and_argument_checks():
    import torch
    from parity_utils import env_params
    from rl_rocket_6dof_b200.batch import Rocket6DOFBatch
    from rl_rocket_6dof_b200.gae import compute_gae
    env = Rocket6DOFBatch(2048, params=env_params(), device="cuda:0", seed=4)
    env.reset()
    traj = env.rollout(160, record=True)
    values = torch.zeros_like(traj["rew"])
    adv, ret = compute_gae(traj["rew"], values, traj["done"], torch.zeros(2048, device="cuda"), 1.0, 1.0)
    torch.cuda.synchronize()
    # zero critic, gamma = lambda = 1: the return at t is the clipped reward-to-go until the episode ends
    rew, done = traj["rew"].cpu().numpy().astype(np.float64), traj["done"].cpu().numpy() != 0
    rtg = np.zeros_like(rew)
    acc = np.zeros(rew.shape[1])
    for t in reversed(range(rew.shape[0])):
        acc = rew[t] + np.where(done[t], 0.0, acc)
        rtg[t] = acc
    assert np.allclose(ret.cpu().numpy(), rtg, rtol=1e-5, atol=1e-4) and done.any()
    with pytest.raises(ValueError):
        compute_gae(traj["rew"], values[:-1], traj["done"], torch.zeros(2048, device="cuda"))
    with pytest.raises(ValueError):
        compute_gae(traj["rew"].double(), values.double(), traj["done"], torch.zeros(2048, device="cuda"))
