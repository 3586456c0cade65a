#!/bin/bash
# round-2 geometry / head kernels: sanitizer passes over tools/geom_probe.py, ncu --set full of FPS + ball query at N = 16384
TAG=${1:-geom}
mkdir -p gpurun_out
timeout 300 python tools/geom_probe.py > gpurun_out/${TAG}_probe.txt 2>&1; echo "probe rc=$?"; tail -n 6 gpurun_out/${TAG}_probe.txt
for tool in racecheck memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/geom_probe.py > gpurun_out/${TAG}_san_$tool.txt 2>&1
  echo "$tool rc=$?"; tail -n 3 gpurun_out/${TAG}_san_$tool.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fps_|ball_query' -c 6 -o gpurun_out/${TAG}_geom16k python bench.py --steps 1 --warmup 1 --no-graph --no-configs --no-cpu-baseline --no-raw-events --points 16384 --windows-per-gpu 256 > gpurun_out/${TAG}_ncu16k.log 2>&1; echo "ncu16k rc=$?"
