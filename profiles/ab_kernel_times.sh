#!/bin/bash
# A/B of library builds in ONE gpurun call: profiles/ab_kernel_times.sh out.jsonl libA.so libB.so ...   (two alternating rounds)
out=$1; shift
: > $out
for rep in 1 2; do for f in "$@"; do
  R6_AUTOBUILD=0 R6_LIB_PATH=$f python profiles/kernel_times.py --tag $(basename $f) >> $out 2>> $out.err
done; done
python - $out <<'PY'
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l)
    print(f"{d['tag']:28s} joined {d['lanes2_joined_ms']*1e3:6.1f}  free {d['lanes2_free_ms']*1e3:6.1f}  1-stream {d['lanes1_ms']*1e3:6.1f} us   {d.get('kernels_us')}")
PY
