#!/bin/bash
# One `ncu --set full` capture of the env-step kernels (lanes = 1 so the launches serialise): profiles/ncu_capture.sh <lib.so> <out.ncu-rep> [kernel regex]
lib=$1; out=$2; rx=${3:-"integrate_first_kernel|integrate_resume_kernel|post_kernel"}
cat > /tmp/ncu_target.py <<'PY'
import torch, sys, os
sys.path.insert(0, os.getcwd())
from rl_rocket_6dof_b200.batch import Rocket6DOFBatch
n = 1 << 20
env = Rocket6DOFBatch(n, device="cuda:0", seed=42, lanes=1, record_attempts=True)
env.reset(); env.rollout(256)
g = torch.Generator(device="cuda"); g.manual_seed(1)
acts = (torch.rand(2, n, 3, device="cuda", generator=g) * 2 - 1).contiguous()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for k in range(2):
    env.step(acts[k])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
PY
R6_AUTOBUILD=0 R6_LIB_PATH=$lib ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k "regex:$rx" -c 4 -f -o ${out%.ncu-rep} python /tmp/ncu_target.py
