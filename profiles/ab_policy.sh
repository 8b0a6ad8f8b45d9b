#!/bin/bash
# A/B of the closed-loop rollout kernels (tc = tensor-core policy, cc = CUDA-core policy) across library builds
for rep in 1 2; do for f in "$@"; do for m in tc cc; do
  echo -n "$(basename $f) "; R6_LIB_PATH=$f python profiles/run_rollout_policy.py $m 1048576 16 | tail -1
done; done; done
