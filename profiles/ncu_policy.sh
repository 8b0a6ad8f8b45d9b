#!/bin/bash
# ncu --set full capture of the policy kernels: profiles/ncu_policy.sh <out.ncu-rep> [tensor_cores mode]
out=$1; mode=${2:-3}
cat > /tmp/ncu_pol.py <<PY
import torch, sys, os
sys.path.insert(0, os.getcwd())
from rl_rocket_6dof_b200 import policy
from rl_rocket_6dof_b200.batch import Rocket6DOFBatch
n = 1 << 20
env = Rocket6DOFBatch(n, device="cuda:0", seed=42)
env.reset(); env.rollout(64)
w = policy.to_device(policy.load_npz("tests/golden/policy_cl.npz"), env.device)
out = torch.empty(n, 3, device="cuda")
for k in range(3): env.policy_actions(w, tensor_cores=$mode, out=out)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
env.policy_actions(w, tensor_cores=$mode, out=out)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
PY
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:policy" -c 1 -f -o ${out%.ncu-rep} python /tmp/ncu_pol.py
