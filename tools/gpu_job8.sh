#!/bin/bash
TAG=${1:-j12}
mkdir -p gpurun_out
EV2H_POOL_LOADER=2 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -x -k "encoder or regressor or compact or sharding or long_window or batch64 or module_by_module" > gpurun_out/${TAG}_pytest_ni1.log 2>&1; echo "pytest ni1 rc=$?"; tail -n 2 gpurun_out/${TAG}_pytest_ni1.log
for v in def pl2; do
  case $v in def) E="";; pl2) E="EV2H_POOL_LOADER=2";; esac
  env $E timeout 300 python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline --no-raw-events > gpurun_out/${TAG}_bench_$v.json 2> gpurun_out/${TAG}_bench_$v.err; echo "bench $v rc=$?"
done
EV2H_POOL_LOADER=2 EV2H_LIB=exp/libev2h_TRACE.so timeout 300 python tools/fused_trace.py tf32x3 > gpurun_out/${TAG}_trace_ni1.txt 2>&1; echo "trace rc=$?"
python - <<PY
import json
for n in ("def","pl2"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_%s.json"%n).read().strip().splitlines()[-1])
        k=d["kernels"]
        print(n, "value %.0f ms %.3f e2e %.0f frac %.3f fused %.3f ms"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], k["ev2h_sa_msg_fused_tc"]["ms_per_step"]))
    except Exception as e: print(n, "failed", e)
PY
grep "====" gpurun_out/${TAG}_trace_ni1.txt
