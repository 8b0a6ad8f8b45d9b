#!/bin/bash
# fused single-kernel r6_step vs integrate|post split, same box, alternating
for rep in 1 2; do for m in 0 1; do
R6_SPLIT_STEP=$m python - <<PY
import torch, os
from rl_rocket_6dof_b200.batch import Rocket6DOFBatch
for prec in ("fp64", "fp32"):
    n=1<<20
    env=Rocket6DOFBatch(n, device="cuda:0", seed=42, precision=prec); env.reset(); env.rollout(256)
    g=torch.Generator(device="cuda"); g.manual_seed(1)
    acts=(torch.rand(8,n,3,device="cuda",generator=g)*2-1)
    for w in range(5): env.step(acts[w%8])
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for k in range(100): env.step(acts[k%8])
    e1.record(); torch.cuda.synchronize()
    print("split=%s %s step %.4f ms" % (os.environ["R6_SPLIT_STEP"], prec, e0.elapsed_time(e1)/100))
PY
done; done
