import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo" if os.path.exists("/root/repo") else ".")
from rl_rocket_6dof_b200 import policy
from rl_rocket_6dof_b200.batch import Rocket6DOFBatch
import ctypes as C
from rl_rocket_6dof_b200 import _lib
from rl_rocket_6dof_b200._lib import R6Mlp
g = np.load("tests/golden/policy_cl.npz")
w = policy.load_npz("tests/golden/policy_cl.npz")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
env = Rocket6DOFBatch(n, device="cuda:0", seed=3)
wd = policy.to_device(w, env.device)
starts = set(int(s) for s in g["ic_step"])
idx = np.array([k for k in range(1, len(g["action"])) if k not in starts])[:n]
env.obs[:, :len(idx)] = torch.from_numpy(np.ascontiguousarray(g["obs"][idx - 1].T)).cuda()
L = _lib.load()
m = _lib.make_mlp(wd)
out = {}
for mode in (0, 2):
    a = torch.zeros(n, 3, device="cuda")
    rc = L.r6_policy(C.byref(m), env.obs.data_ptr(), n, mode, a.data_ptr(), None)
    print("mode", mode, "rc", rc, L.r6_last_error())
    torch.cuda.synchronize()
    out[mode] = a.cpu().numpy()
d = np.abs(out[2] - out[0])
print("max |d action| tcgen05 vs fp32:", d.max(), "mean", d.mean())
print(out[0][:3], out[2][:3])
