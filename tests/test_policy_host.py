"""Fused policy network (host build of the kernel code) against the actions the reference-side
closed-loop run recorded (tests/golden/policy_cl.npz: numpy float32 forward of best_model_2bo71j9m)."""
import os

import numpy as np

import hostsim
from rl_rocket_6dof_b200 import policy

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "policy_cl.npz")


def _episode_inputs(g):
    """(obs13 the policy saw, action it produced) for every step that is not the first of an episode."""
    starts = set(int(s) for s in g["ic_step"])
    idx = np.array([k for k in range(1, len(g["action"])) if k not in starts])
    return g["obs"][idx - 1][:, :13], g["action"][idx]


def test_weights_roundtrip_and_shapes():
    w = policy.load_npz(GOLD)
    assert {k: v.shape for k, v in w.items()} == policy.SHAPES
    bad = dict(w, w0=w["w0"][:, :12])
    try:
        policy._check(bad)
    except ValueError as e:
        assert "w0" in str(e)
    else:
        raise AssertionError("wrong shapes must be rejected")


def test_kernel_mlp_matches_recorded_actions():
    g = np.load(GOLD)
    w = policy.load_npz(GOLD)
    x, a_ref = _episode_inputs(g)
    a = hostsim.mlp_actions(w, x)
    # float32 network: summation order / tanh implementation differ by a few ulp of the pre-activations
    assert np.abs(a - a_ref).max() <= 2e-6
    assert np.abs(policy.forward_numpy(w, x) - a_ref).max() <= 2e-6
    assert np.all(np.abs(a) <= 1.0)
    # the clip of evaluate_policy / predict: a network with a large output gain saturates at +-1
    big = dict(w, w2=w["w2"] * 1000, b2=w["b2"] * 1000)
    ab = hostsim.mlp_actions(big, x[:256])
    assert np.array_equal(ab, policy.forward_numpy(big, x[:256])) or np.abs(ab - policy.forward_numpy(big, x[:256])).max() <= 1e-3
    assert (np.abs(ab) == 1.0).mean() > 0.5


def test_sb3_zip_loader_roundtrip(tmp_path):
    """policy.load_sb3_zip reads the actor (+ critic head and log_std of the shared-trunk policy) out of a
    stable-baselines3 archive layout: a zip with a `policy.pth` state dict."""
    import io
    import zipfile

    import torch
    w = policy.load_npz(GOLD)
    rng = np.random.default_rng(2)
    sd = {
        "mlp_extractor.shared_net.0.weight": torch.from_numpy(w["w0"]), "mlp_extractor.shared_net.0.bias": torch.from_numpy(w["b0"]),
        "mlp_extractor.shared_net.2.weight": torch.from_numpy(w["w1"]), "mlp_extractor.shared_net.2.bias": torch.from_numpy(w["b1"]),
        "action_net.weight": torch.from_numpy(w["w2"]), "action_net.bias": torch.from_numpy(w["b2"]),
        "value_net.weight": torch.from_numpy(rng.standard_normal((1, 64)).astype(np.float32)),
        "value_net.bias": torch.zeros(1), "log_std": torch.full((3,), -1.5),
    }
    buf = io.BytesIO()
    torch.save(sd, buf)
    path = str(tmp_path / "model.zip")
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("policy.pth", buf.getvalue())
        z.writestr("data", "{}")
    got = policy.load_sb3_zip(path[:-4])                 # extension is optional, like PPO.load
    for k in policy.SHAPES:
        assert np.array_equal(got[k], w[k])
    assert got["wv"].shape == (64,) and got["bv"].shape == (1,) and np.allclose(got["log_std"], -1.5)
    mean, value = policy.forward_full_numpy(got, np.zeros((2, 13), np.float32))
    assert mean.shape == (2, 3) and value.shape == (2,)
    a = hostsim.mlp_actions(got, np.zeros((2, 13), np.float32))
    assert np.abs(a - np.clip(mean, -1, 1)).max() <= 1e-6
