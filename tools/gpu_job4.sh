#!/bin/bash
TAG=${1:-j7}
mkdir -p gpurun_out
EV2H_FUSED_OCC=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline --no-raw-events > gpurun_out/${TAG}_bench_occ1.json 2> gpurun_out/${TAG}_bench_occ1.err; echo "bench occ1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sa_fused -s 15 -c 5 -o gpurun_out/${TAG}_fused python bench.py --steps 1 --warmup 3 --no-graph --no-configs --no-cpu-baseline --no-raw-events > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
python - <<PY
import json
for n in ("occ1",):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_%s.json"%n).read().strip().splitlines()[-1])
        k=d["kernels"]
        print(n, "value %.0f ms %.3f e2e %.0f frac %.3f fused %.3f ms"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], k["ev2h_sa_msg_fused_tc"]["ms_per_step"]))
    except Exception as e: print(n, "failed", e)
PY
