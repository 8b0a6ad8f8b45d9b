import time, numpy as np, torch, sys
sys.path.insert(0, ".")
from rl_rocket_6dof_b200 import make_vec_env
for n in (1, 64, 4096, 65536):
    env = make_vec_env(n, device="cuda:0", seed=1)
    env.reset()
    a = np.random.default_rng(0).uniform(-1, 1, (n, 3)).astype(np.float32)
    for _ in range(20): env.step(a)
    t0 = time.perf_counter(); K = 200
    for _ in range(K): env.step(a)
    t1 = time.perf_counter()
    for _ in range(K): env.step_host(a)
    t2 = time.perf_counter()
    print(f"n={n:6d}: VecEnv.step {1e6*(t1-t0)/K:8.1f} us ({n*K/(t1-t0):.3e} env-steps/s) | step_host {1e6*(t2-t1)/K:8.1f} us ({n*K/(t2-t1):.3e})")
