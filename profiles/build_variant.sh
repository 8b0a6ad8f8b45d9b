#!/bin/bash
# Builds rl_rocket_6dof_b200/lib/libr6dof_<tag>.so with extra -D flags for A/B timing (R6_LIB_PATH=<that .so>).
# Usage: profiles/build_variant.sh <tag> [-DFLAG=V ...]      prints registers / spills of the hot kernels
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
out=rl_rocket_6dof_b200/lib/libr6dof_$tag.so
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC -shared \
     "$@" -Xptxas -v -o $out rl_rocket_6dof_b200/csrc/r6_kernels.cu > rl_rocket_6dof_b200/lib/build_$tag.log 2>&1
python - "$tag" <<'PY'
import re, sys
log = open(f"rl_rocket_6dof_b200/lib/build_{sys.argv[1]}.log").read()
for m in re.finditer(r"Compiling entry function '(\S+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", log):
    name = m.group(1)
    for k in ("integrate_first_kernelIdLb0", "integrate_resume_kernelIdLb0ELb0", "integrate_resume_kernelIdLb0ELb1", "integrate_kernelIdLb0", "post_kernelId", "post_pipe_kernelId", "step_kernelIdLb0", "tail_kernelId"):
        if k in name:
            print(f"  {k:32s} regs {m.group(5):>3s}  stack {m.group(2):>4s}  spill st/ld {m.group(3)}/{m.group(4)}")
PY
echo "-> $out"
