#!/bin/bash
# Builds libr6dof_t<threads>_b<minblocks>.so variants (CTA size / resident-CTA target) for A/B timing with
# profiles/sweep_variants.sh.  Usage: profiles/build_variants.sh "128 3" "128 4" "256 2" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p rl_rocket_6dof_b200/lib
for v in "$@"; do
  set -- $v
  out=rl_rocket_6dof_b200/lib/libr6dof_t$1_b$2.so
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC -shared \
       -DR6_THREADS=$1 -DR6_MIN_BLOCKS=$2 -Xptxas -v -o $out rl_rocket_6dof_b200/csrc/r6_kernels.cu 2>&1 |
    grep -A2 "step_kernelILb0" | grep -E "registers|spill" | tr '\n' ' '
  echo " -> $out"
done
