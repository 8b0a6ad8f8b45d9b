#!/bin/bash
# quick A/B of the step kernel only (same box, alternating): profiles/ab_step.sh libA.so libB.so ...
for rep in 1 2; do
for f in "$@"; do
  R6_LIB_PATH=$f python - <<PY
import torch, os
from rl_rocket_6dof_b200.batch import Rocket6DOFBatch
n=1<<20
env=Rocket6DOFBatch(n, device="cuda:0", seed=42); env.reset(); env.rollout(256)
g=torch.Generator(device="cuda"); g.manual_seed(1)
acts=(torch.rand(8,n,3,device="cuda",generator=g)*2-1)
for w in range(5): env.step(acts[w%8])
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for k in range(100): env.step(acts[k%8])
e1.record(); torch.cuda.synchronize()
print(os.environ["R6_LIB_PATH"].split("/")[-1], "step %.4f ms" % (e0.elapsed_time(e1)/100))
PY
done; done
