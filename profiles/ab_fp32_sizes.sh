#!/bin/bash
for f in "$@"; do
R6_LIB_PATH=$f python - <<PY
import torch, os
from rl_rocket_6dof_b200.batch import Rocket6DOFBatch
for prec in ("fp32","fp64"):
  for lg in (20, 23):
    n=1<<lg
    env=Rocket6DOFBatch(n, device="cuda:0", seed=42, precision=prec); env.reset(); env.rollout(256)
    g=torch.Generator(device="cuda"); g.manual_seed(1)
    acts=(torch.rand(4,n,3,device="cuda",generator=g)*2-1)
    for w in range(5): env.step(acts[w%4])
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    k = 100 if lg==20 else 20
    for j in range(k): env.step(acts[j%4])
    e1.record(); torch.cuda.synchronize()
    print(os.environ["R6_LIB_PATH"].split("/")[-1], prec, "2^%d"%lg, "%.4f ms/step  %.3e env-steps/s" % (e0.elapsed_time(e1)/k, n*k/(e0.elapsed_time(e1)*1e-3)))
    del env
PY
done
