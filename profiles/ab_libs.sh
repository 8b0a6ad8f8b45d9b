#!/bin/bash
# A/B timing of library builds in ONE gpurun call (same box, alternating): profiles/ab_libs.sh libA.so libB.so ...
for rep in 1 2; do
for f in "$@"; do
  R6_LIB_PATH=$f python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print('$f', 'step %.4f ms | rollout %.4f ms | fp32 %.4f ms | policy %.4f ms | e2e %.3f ms' % (d['ms_per_step'], d['rollout_fused']['ms_per_step'], d['fp32_path']['ms_per_step'], d['rollout_policy']['ms_per_step'], d['e2e']['ms_per_step']))"
done; done
