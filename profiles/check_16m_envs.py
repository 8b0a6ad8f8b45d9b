import torch, sys
sys.path.insert(0, ".")
from rl_rocket_6dof_b200.batch import Rocket6DOFBatch
n = 1 << 24
env = Rocket6DOFBatch(n, device="cuda:0", seed=5)
env.reset()
env.step_random(40)
a = torch.rand(n, 3, device="cuda") * 2 - 1
env.step(a)
torch.cuda.synchronize()
s = env.stats.cpu().numpy()
print("envs", n, "steps", s[7], "expected", n * 41, "episodes", s[0], "finite", bool(torch.isfinite(env.state).all()), bool(torch.isfinite(env.obs).all()))
q = env.state[6:10]
print("quat norm err", float((q.pow(2).sum(0).sqrt() - 1).abs().max()), "step_count max", int(env.step_count.max()), "last env state", env.state[:3, -1].tolist())
print("mem GB", torch.cuda.max_memory_allocated() / 1e9)
