#!/bin/bash
# sanitizer passes over the fused kernel (small problem), ncu of the geometry kernels at N = 16384, launch list of one step
TAG=${1:-j11}
mkdir -p gpurun_out
for tool in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/fused_probe.py a128 b > gpurun_out/${TAG}_san_$tool.txt 2>&1
  echo "$tool rc=$?"; tail -n 4 gpurun_out/${TAG}_san_$tool.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fps_|ball_query' -c 6 -o gpurun_out/${TAG}_geom16k python bench.py --steps 1 --warmup 1 --no-graph --no-configs --no-cpu-baseline --no-raw-events --points 16384 --windows-per-gpu 256 > gpurun_out/${TAG}_ncu16k.log 2>&1; echo "ncu16k rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-configs --no-cpu-baseline --no-raw-events > gpurun_out/${TAG}_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sa_fused -s 15 -c 5 -o gpurun_out/${TAG}_fused python bench.py --steps 1 --warmup 3 --no-graph --no-configs --no-cpu-baseline --no-raw-events > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu fused rc=$?"
