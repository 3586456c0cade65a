#!/bin/bash
# A/B of builds on ONE box, interleaved: the in-tree library against exp/libev2h_<name>.so variants (tools/build_variant.sh)
TAG=${1:-ab}; shift
mkdir -p gpurun_out
for rep in a b; do
  for v in tree "$@"; do
    if [ $v = tree ]; then E="X=0"; else E="EV2H_LIB=$PWD/exp/libev2h_$v.so"; fi
    env $E timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline --no-raw-events 2> /dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); k = d['kernels']
print('$v $rep: %.0f windows/s %.3f ms/step, fused %.3f ms' % (d['value'], d['ms_per_step'], k['ev2h_sa_msg_fused_tc']['ms_per_step']))" | tee -a gpurun_out/${TAG}_ab.txt
  done
done
