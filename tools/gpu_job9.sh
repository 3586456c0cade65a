#!/bin/bash
TAG=${1:-j15}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -x -k "encoder or regressor or compact or sharding or long_window or batch64 or module_by_module or graph or data_parallel or tehnet" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/${TAG}_pytest.log
for v in def noss; do
  case $v in def) E="";; noss) E="EV2H_SCALE_STREAMS=0";; esac
  env $E timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline --no-raw-events > gpurun_out/${TAG}_bench_$v.json 2> gpurun_out/${TAG}_bench_$v.err; echo "bench $v rc=$?"
done
python - <<PY
import json
for n in ("def","noss"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_%s.json"%n).read().strip().splitlines()[-1])
        k=d["kernels"]
        print(n, "value %.0f ms %.3f e2e %.0f frac %.3f fused %.3f ms"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], k["ev2h_sa_msg_fused_tc"]["ms_per_step"]), d["config"]["launch"])
    except Exception as e: print(n, "failed", e)
PY
