"""Edge cases of the C-ABI path on the GPU: empty and ragged batches, masked resets, zero-step rollouts,
argument errors, the out-of-range density fallback and a NaN-poisoned env that must not disturb its neighbours."""
import ctypes as C

import numpy as np
import pytest

from parity_utils import env_params

pytestmark = pytest.mark.gpu


def _mk(n, **kw):
    from rl_rocket_6dof_b200.batch import Rocket6DOFBatch
    kw.setdefault("params", env_params())
    return Rocket6DOFBatch(n, device="cuda:0", seed=kw.pop("seed", 1), **kw)


@pytest.mark.parametrize("n", [1, 31, 129, 1000])
def test_ragged_batches_equal_slices_of_a_big_batch(n):
    """Envs are independent and Philox is keyed by global env id: a batch of n envs is bit-identical to the
    first n envs of a larger batch, whatever the CTA / warp padding."""
    import torch
    big, small = _mk(4096, seed=3), _mk(n, seed=3, num_envs_global=4096)
    big.reset(); small.reset()
    big.rollout(40); small.rollout(40)
    torch.cuda.synchronize()
    assert torch.equal(big.state[:, :n], small.state) and torch.equal(big.obs[:, :n], small.obs)
    assert torch.equal(big.episode_id[:n], small.episode_id) and torch.equal(big.reward[:n], small.reward)
    assert float(small.stats[7]) == n * 40


def test_empty_batch_and_zero_step_rollout_are_noops():
    import torch
    env = _mk(64)
    env.reset()
    snap = env.state.clone()
    L = env.lib
    assert L.r6_step(C.byref(env._p), C.byref(env._b), 0, 0, env.obs.data_ptr(), 0, None) == 0          # n = 0
    assert L.r6_reset(C.byref(env._p), C.byref(env._b), 0, 0, None, 0, None) == 0
    env.rollout(0)
    torch.cuda.synchronize()
    assert torch.equal(env.state, snap) and float(env.stats[7]) == 0
    assert L.r6_gae(env.obs.data_ptr(), env.obs.data_ptr(), env.done.data_ptr(), env.obs.data_ptr(), 0, 64, 0.99, 0.95,
                    env.obs.data_ptr(), env.obs.data_ptr(), None) == 0


def test_argument_errors_have_codes_and_messages():
    env = _mk(8)
    L = env.lib
    assert L.r6_step(C.byref(env._p), C.byref(env._b), -1, 0, env.obs.data_ptr(), 0, None) == -1
    assert b"n < 0" in L.r6_last_error()
    assert L.r6_step(C.byref(env._p), C.byref(env._b), 8, 0, None, 0, None) == -1 and b"actions" in L.r6_last_error()
    assert L.r6_rollout(C.byref(env._p), C.byref(env._b), 8, 0, 4, 99, None, None, 0, 0, None, None, None, None, None) == -1
    assert b"action mode" in L.r6_last_error()
    assert L.r6_rollout(C.byref(env._p), C.byref(env._b), 8, 0, 4, 2, None, None, 0, 0, None, None, None, None, None) == -1
    assert b"act_buf" in L.r6_last_error()
    p = env.params.to_struct()
    p.precision = 7
    assert L.r6_reset(C.byref(p), C.byref(env._b), 8, 0, None, 0, None) == -1 and b"precision" in L.r6_last_error()
    with pytest.raises(ValueError):
        env.step(env.obs[:3].t().contiguous().double())
    with pytest.raises(ValueError):
        _mk(8, precision="fp16")


def test_masked_reset_only_touches_selected_envs():
    import torch
    env = _mk(300, auto_reset=False)
    env.reset()
    env.rollout(30)
    before = env.state.clone()
    ep_before = env.episode_id.clone()
    mask = torch.zeros(300, dtype=torch.uint8, device="cuda")
    mask[::7] = 1
    env.reset(mask)
    torch.cuda.synchronize()
    m = mask.bool()
    assert torch.equal(env.state[:, ~m], before[:, ~m]) and torch.equal(env.episode_id[~m], ep_before[~m])
    assert bool((env.step_count[m] == 0).all()) and bool((env.episode_id[m] == ep_before[m] + 1).all())
    assert bool((env.step_count[~m] == 30).all())
    ic_lo = torch.as_tensor(env.params.ic_low, device="cuda")[:, None]
    ic_hi = torch.as_tensor(env.params.ic_high, device="cuda")[:, None]
    cols = [0, 1, 2, 3, 4, 5, 13]
    s32 = env.state[:, m].float()
    assert bool(((s32[cols] >= ic_lo[cols]) & (s32[cols] <= ic_hi[cols])).all())


def test_density_fallback_outside_series_range():
    """Initial heights beyond the range of the binomial series (|c h| > 0.08, h > 3.5 km) take the out-of-line pow:
    r6_sim_step_raw at 20 km against the oracle."""
    import torch
    from oracle import c_oracle as co
    from rl_rocket_6dof_b200 import _lib
    L = _lib.load()
    n = 64
    rng = np.random.default_rng(0)
    y = np.zeros((n, 14))
    y[:, 0] = rng.uniform(4000, 20000, n); y[:, 3] = rng.uniform(-300, -50, n)
    q = rng.normal(0, 0.05, (n, 4)); q[:, 0] = 1; y[:, 6:10] = q / np.linalg.norm(q, axis=1)[:, None]
    y[:, 10:13] = rng.normal(0, 0.02, (n, 3)); y[:, 13] = 45e3
    u = np.stack([rng.uniform(-0.2, 0.2, n), rng.uniform(-0.2, 0.2, n), rng.uniform(2e5, 9e5, n)], 1)
    st = torch.from_numpy(np.ascontiguousarray(y.T)).cuda()
    ud = torch.from_numpy(np.ascontiguousarray(u.T)).cuda()
    m0 = torch.full((n,), 45e3, dtype=torch.float64, device="cuda")
    t = torch.zeros(n, dtype=torch.float64, device="cuda")
    status = torch.zeros(n, dtype=torch.int8, device="cuda")
    natt = torch.zeros(n, dtype=torch.uint8, device="cuda")
    _lib.check(L.r6_sim_step_raw(st.data_ptr(), ud.data_ptr(), m0.data_ptr(), t.data_ptr(), 0.1, n, status.data_ptr(),
                                 natt.data_ptr(), None), L)
    torch.cuda.synchronize()
    out = st.t().cpu().numpy()
    for i in range(n):
        ref, rst, nfev = co.sim_step_raw(y[i], u[i], 45e3, 0.0, 0.1)
        assert rst == int(status[i]) and nfev == 2 + 6 * int(natt[i])
        assert np.max(np.abs(out[i] - ref) / np.maximum(np.abs(ref), 1e-3)) <= 1e-9


def test_nan_env_does_not_disturb_neighbours():
    import torch
    a, b = _mk(256, seed=8), _mk(256, seed=8)
    a.reset(); b.reset()
    b.state[3, 100] = float("nan")                       # poison one env's velocity
    acts = torch.rand(20, 256, 3, device="cuda") * 2 - 1
    for j in range(20):
        a.step(acts[j]); b.step(acts[j])
    torch.cuda.synchronize()
    keep = torch.ones(256, dtype=torch.bool, device="cuda"); keep[100] = False
    assert torch.equal(a.state[:, keep], b.state[:, keep]) and torch.equal(a.reward[keep], b.reward[keep])


@pytest.mark.parametrize("split", [False, True])
def test_no_tgo_root_keeps_reward_and_statistics_finite(split):
    """r = 0 exactly: the t_go quartic g^2 t^4 - 4 |v|^2 t^2 has no constant term and the reference raises IndexError
    (rocket_env.py:545).  The kernels define the target acceleration as that of t_go -> infinity instead, so the reward
    is finite and one such env cannot poison the batch statistics (a NaN return went into the shared sums)."""
    import torch
    n = 300
    env = _mk(n, auto_reset=True, split_step=split, seed=4)
    env.reset()
    ic = env.state.t().clone().to(torch.float32)
    ic[7, 0:3] = 0.0                                     # env 7 sits exactly at the origin, moving
    env.set_state(ic[7:8], torch.tensor([7]))
    acts = torch.zeros(n, 3, device="cuda")
    env.step(acts)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(env.reward).all()) and bool(torch.isfinite(env.ep_return).all())
    env.rollout(400)                                      # every env finishes at least one episode
    torch.cuda.synchronize()
    s = env.stats.cpu().numpy()
    assert np.all(np.isfinite(s)) and s[0] >= n


def test_step_random_equals_fused_rollout():
    """r6_step_random (integrator | post-step kernels, in-kernel Philox) == r6_rollout(R6_ACT_PHILOX), step by step."""
    import torch
    a, b = _mk(3000, seed=12), _mk(3000, seed=12)
    a.reset(); b.reset()
    a.step_random(90)
    b.rollout(90)
    torch.cuda.synchronize()
    assert torch.equal(a.episode_id, b.episode_id) and torch.equal(a.step_count, b.step_count)
    norm = torch.as_tensor(a.params.state_normalizer, device="cuda")[:, None]
    assert float(((a.state - b.state).abs() / norm).max()) <= 1e-11       # different kernels: FMA contraction may differ
    sa, sb = a.stats.cpu().numpy(), b.stats.cpu().numpy()
    assert np.array_equal(sa[[0, 2, 3, 4, 5, 6, 7]], sb[[0, 2, 3, 4, 5, 6, 7]]) and a.steps_done == b.steps_done == 90


def test_exact_density_variants_of_the_rollout_kernels():
    """dt = 0.5 s selects the kExact instantiations (true pow density): fused Philox rollout == step-by-step with the
    same actions, and the tensor-core policy request falls back to the float32 network (no exact tensor-core kernel)."""
    import torch
    import philox_ref as pr
    from rl_rocket_6dof_b200 import policy
    from rl_rocket_6dof_b200.batch import ACT_MLP, ACT_MLP_TC
    import os
    ep = env_params(timestep=0.5)
    n, K, seed = 700, 25, 99
    a, b = _mk(n, params=ep, seed=seed), _mk(n, params=ep, seed=seed)
    a.reset(); b.reset()
    a.rollout(K, fused=True)
    for j in range(K):
        b.step(torch.from_numpy(pr.actions(seed, np.arange(n), j)).cuda())
    torch.cuda.synchronize()
    norm = torch.as_tensor(ep.state_normalizer, device="cuda")[:, None]
    assert torch.equal(a.episode_id, b.episode_id) and float(((a.state - b.state).abs() / norm).max()) <= 1e-11
    w = policy.to_device(policy.load_npz(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "policy_cl.npz")), "cuda:0")
    c, d = _mk(n, params=ep, seed=seed), _mk(n, params=ep, seed=seed)
    c.reset(); d.reset()
    c.rollout(K, ACT_MLP, mlp=w); d.rollout(K, ACT_MLP_TC, mlp=w)
    torch.cuda.synchronize()
    assert torch.equal(c.state, d.state)
