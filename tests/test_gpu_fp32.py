"""The optional float32 throughput path (R6_PREC_F32) against the float64 parity path on the same
initial states and action sequences.  Its bound is its own (north_star): the float32 integrator
carries ~6e-8 relative round-off per operation and the dynamics amplify it 10-100x over an episode
(SURVEY §8d), so after 100 env-steps the state agrees to <= 2e-3 of the normaliser (measured:
printed below), episode endings are identical except for envs that sit within float32 round-off
of a threshold, and the reward agrees to <= 2e-3 absolute except on the (<= 1e-4 of all) steps where
a discrete reward term (attitude-limit penalty, goal bonus) flips because the state sits within
float32 round-off of its threshold."""
import numpy as np
import pytest

from parity_utils import env_params

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("multipass", [False, True])
def test_fp32_path_tracks_fp64_path(multipass):
    """multipass=True: the dispatch large float32 batches take by default (one RK attempt per pass, two stream lanes)."""
    import torch
    from rl_rocket_6dof_b200.batch import Rocket6DOFBatch
    ep = env_params()
    n, k = 8192, 100
    kw = dict(params=ep, device="cuda:0", auto_reset=False, clip_reward=False, time_limit=False, seed=21,
              record_attempts=True)
    a = Rocket6DOFBatch(n, precision="fp64", **kw)
    b = Rocket6DOFBatch(n, precision="fp32", **(dict(kw, split_step=True, multipass=True, lanes=2) if multipass else kw))
    a.reset(); b.reset()
    assert b.state.dtype == torch.float32 and b.terminal_state.dtype == torch.float32
    assert torch.equal(a.state.to(torch.float32), b.state)          # same Philox initial conditions
    norm = torch.as_tensor(ep.state_normalizer, device="cuda")[:, None]
    gen = torch.Generator(device="cuda"); gen.manual_seed(5)
    alive = torch.ones(n, dtype=torch.bool, device="cuda")
    worst = torch.zeros(14, device="cuda", dtype=torch.float64)
    worst_rew, n_att_diff, n_alive_steps, split, n_rew_flip = 0.0, 0, 0, 0, 0
    for j in range(k):
        act = (torch.rand(n, 3, device="cuda", generator=gen) * 2 - 1)
        a.step(act); b.step(act)
        da, db = (a.flags & 3) != 0, (b.flags & 3) != 0
        err = ((a.state - b.state.to(torch.float64)).abs() / norm)[:, alive]
        worst = torch.maximum(worst, err.max(dim=1).values)
        dr = (a.reward - b.reward).abs()[alive & ~da & ~db]
        n_rew_flip += int((dr > 2e-3).sum())
        worst_rew = max(worst_rew, float(dr[dr <= 2e-3].max()))
        n_att_diff += int((a.nattempts != b.nattempts)[alive].sum())
        n_alive_steps += int(alive.sum())
        split += int((da != db)[alive].sum())
        alive &= ~(da | db)                  # compare episodes only while both are running
    torch.cuda.synchronize()
    print("fp32 vs fp64 after <=100 steps: max |dx|/normaliser per component:", worst.cpu().numpy())
    print(f"max |d reward| {worst_rew:.3e} (discrete-term flips: {n_rew_flip}); RK-attempt mismatches {n_att_diff}/{n_alive_steps}; "
          f"episode endings that differ by a step {split}/{n}")
    assert float(worst.max()) <= 2e-3
    assert worst_rew <= 2e-3 and n_rew_flip <= 1e-4 * n_alive_steps
    assert n_att_diff <= 2e-3 * n_alive_steps
    assert split <= 0.01 * n


def test_fp32_autoreset_rollout_statistics_match_fp64():
    """Distributional check over whole episodes: same Philox ICs / actions, 400 fused steps."""
    import torch
    from rl_rocket_6dof_b200.batch import Rocket6DOFBatch
    ep = env_params()
    n, k = 16384, 400
    out = {}
    for prec in ("fp64", "fp32"):
        env = Rocket6DOFBatch(n, params=ep, device="cuda:0", seed=77, precision=prec)
        env.reset()
        env.rollout(k)
        torch.cuda.synchronize()
        out[prec] = env.stats_dict()
    a, b = out["fp64"], out["fp32"]
    print("fp64:", a, "\nfp32:", b)
    assert a["steps"] == b["steps"] == n * k
    assert abs(a["episodes"] - b["episodes"]) <= 0.002 * a["episodes"]
    assert abs(a["mean_length"] - b["mean_length"]) <= 0.002 * a["mean_length"]
    assert abs(a["mean_return"] - b["mean_return"]) <= 0.01 * abs(a["mean_return"])
    assert abs(a["out_of_bounds"] - b["out_of_bounds"]) <= 0.005 * a["episodes"]


def test_fp32_split_kernels_equal_fused_kernel():
    """float32 path: r6_step as the integrator | post-step kernel pair == the fused kernel (same per-env code; the
    two builds may contract FMAs differently, so float32 round-off, identical episode boundaries)."""
    import torch
    from rl_rocket_6dof_b200.batch import Rocket6DOFBatch
    ep = env_params()
    n, k = 4096, 120
    a = Rocket6DOFBatch(n, params=ep, device="cuda:0", seed=9, precision="fp32", split_step=True)
    b = Rocket6DOFBatch(n, params=ep, device="cuda:0", seed=9, precision="fp32", split_step=False)
    assert a.scratch is not None and b.scratch is None
    a.reset(); b.reset()
    gen = torch.Generator(device="cuda"); gen.manual_seed(2)
    norm = torch.as_tensor(ep.state_normalizer, device="cuda", dtype=torch.float32)[:, None]
    for j in range(k):
        act = torch.rand(n, 3, device="cuda", generator=gen) * 2 - 1
        a.step(act); b.step(act)
    torch.cuda.synchronize()
    same = (a.episode_id == b.episode_id) & (a.step_count == b.step_count)
    assert float(same.float().mean()) >= 0.999                 # float32 round-off may move an episode end by a step
    assert float(((a.state - b.state).abs() / norm)[:, same].max()) <= 1e-4
    assert abs(float(a.stats[0]) - float(b.stats[0])) <= 2
