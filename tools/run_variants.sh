#!/bin/bash
# On the GPU box: time the encoder bench with each experimental build in exp/ (EV2H_LIB override).
TAG=${1:-var}; shift
for V in "$@"; do
  for P in ${MLPS:-tf32x3 bf16}; do
    R=$(EV2H_LIB=$PWD/exp/libev2h_$V.so timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --mlp $P 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); k = d['kernels']
print('%.3f ms/step, fused %.3f ms' % (d['ms_per_step'], k['ev2h_sa_msg_fused_tc']['ms_per_step']))")
    echo "$V $P: $R" | tee -a gpurun_out/${TAG}_variants.txt
  done
done
