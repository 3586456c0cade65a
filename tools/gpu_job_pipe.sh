#!/bin/bash
TAG=${1:-pipe}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/${TAG}_pytest.log
run() {  # label, env...
  L=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline --no-raw-events 2> gpurun_out/${TAG}_err.txt | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); k = d['kernels']
print('$L: %.0f windows/s %.3f ms/step e2e %.0f, fps %.3f ms ball %.3f fused %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], k['ev2h_fps_f32']['ms_per_step'], k['ev2h_ball_query_f32']['ms_per_step'], k['ev2h_sa_msg_fused_tc']['ms_per_step']))" | tee -a gpurun_out/${TAG}_ab.txt
}
for rep in a b; do
  run "splits4 $rep" EV2H_FPS_SPLITS=4
  run "splits0 $rep" EV2H_FPS_SPLITS=0
  run "splits2 $rep" EV2H_FPS_SPLITS=2
  run "splits8 $rep" EV2H_FPS_SPLITS=8
  run "scan+splits4 $rep" EV2H_LIB=$PWD/exp/libev2h_fpsscan.so EV2H_FPS_SPLITS=4
  run "t128+splits0 $rep" EV2H_FPS_THREADS=128 EV2H_FPS_SPLITS=0
done
