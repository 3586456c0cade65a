#!/bin/bash
# iteration job: default build vs EV2H_FUSED_LG2=1 (two loader groups at two CTAs per SM) and EV2H_FUSED_OCC=1
TAG=${1:-j8}
mkdir -p gpurun_out
SUB="encoder or regressor or compact or sharding or long_window or batch64 or module_by_module"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -x -k "$SUB" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
EV2H_FUSED_LG2=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -x -k "$SUB" > gpurun_out/${TAG}_pytest_lg2.log 2>&1; echo "pytest lg2 rc=$?"
EV2H_FUSED_OCC=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "encoder_golden or compact" > gpurun_out/${TAG}_pytest_occ1.log 2>&1; echo "pytest occ1 rc=$?"
for v in def lg2 occ1; do
  case $v in def) E="";; lg2) E="EV2H_FUSED_LG2=1";; occ1) E="EV2H_FUSED_OCC=1";; esac
  env $E timeout 300 python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline --no-raw-events > gpurun_out/${TAG}_bench_$v.json 2> gpurun_out/${TAG}_bench_$v.err; echo "bench $v rc=$?"
done
EV2H_FUSED_LG2=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline --no-raw-events --mlp bf16 > gpurun_out/${TAG}_bench_lg2bf16.json 2> gpurun_out/${TAG}_bench_lg2bf16.err
EV2H_FUSED_LG2=1 EV2H_LIB=exp/libev2h_TRACE.so timeout 300 python tools/fused_trace.py tf32x3 > gpurun_out/${TAG}_trace_lg2.txt 2>&1; echo "trace rc=$?"
python - <<PY
import json
for n in ("def","lg2","occ1","lg2bf16"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_%s.json"%n).read().strip().splitlines()[-1])
        k=d["kernels"]
        print(n, "value %.0f ms %.3f e2e %.0f frac %.3f fused %.3f ms"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], k["ev2h_sa_msg_fused_tc"]["ms_per_step"]))
    except Exception as e: print(n, "failed", e)
PY
grep "====" gpurun_out/${TAG}_trace_lg2.txt
tail -2 gpurun_out/${TAG}_pytest.log gpurun_out/${TAG}_pytest_lg2.log gpurun_out/${TAG}_pytest_occ1.log
