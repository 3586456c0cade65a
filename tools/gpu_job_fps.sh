#!/bin/bash
TAG=${1:-fps}
mkdir -p gpurun_out
EV2H_LIB=$PWD/exp/libev2h_fpsscan.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -x -k "fps or encoder_golden or long_window" > gpurun_out/${TAG}_pytest_scan.log 2>&1; echo "pytest scan rc=$?"; tail -n 2 gpurun_out/${TAG}_pytest_scan.log
EV2H_FPS_THREADS=128 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fps or encoder_golden" > gpurun_out/${TAG}_pytest_t128.log 2>&1; echo "pytest t128 rc=$?"; tail -n 2 gpurun_out/${TAG}_pytest_t128.log
run() {  # label, env...
  L=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline --no-raw-events 2> /dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); k = d['kernels']
print('$L: %.0f windows/s %.3f ms/step, fps %.3f ms ball %.3f fused %.3f' % (d['value'], d['ms_per_step'], k['ev2h_fps_f32']['ms_per_step'], k['ev2h_ball_query_f32']['ms_per_step'], k['ev2h_sa_msg_fused_tc']['ms_per_step']))" | tee -a gpurun_out/${TAG}_ab.txt
}
for rep in a b; do
  run "tree $rep" X=0
  run "scan $rep" EV2H_LIB=$PWD/exp/libev2h_fpsscan.so
  run "t128 $rep" EV2H_FPS_THREADS=128
  run "t512 $rep" EV2H_FPS_THREADS=512
  run "scan+t128 $rep" EV2H_LIB=$PWD/exp/libev2h_fpsscan.so EV2H_FPS_THREADS=128
done
