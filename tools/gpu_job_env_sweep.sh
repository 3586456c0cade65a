#!/bin/bash
# sweep of environment switches on ONE box with the in-tree library (interleaved, two rounds)
TAG=${1:-sw}; shift
mkdir -p gpurun_out
for rep in a b; do
  for E in "$@"; do
    env $(echo $E | tr ',' ' ') timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline --no-raw-events 2> /dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); k = d['kernels']
print('$E $rep: %.0f windows/s %.3f ms/step, fused %.3f ms' % (d['value'], d['ms_per_step'], k['ev2h_sa_msg_fused_tc']['ms_per_step']))" | tee -a gpurun_out/${TAG}_sweep.txt
  done
done
