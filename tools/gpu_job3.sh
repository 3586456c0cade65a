#!/bin/bash
# quick iteration job: parity subset, headline bench (fp32-level + bf16), timeline trace
TAG=${1:-j3}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -x -k "encoder or regressor or compact or sharding or linear_relu or long_window or batch64 or module_by_module or decoder" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline --no-raw-events > gpurun_out/${TAG}_bench_f16.json 2> gpurun_out/${TAG}_bench_f16.err; echo "bench f16 rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline --no-raw-events --mlp bf16 > gpurun_out/${TAG}_bench_bf16.json 2> gpurun_out/${TAG}_bench_bf16.err; echo "bench bf16 rc=$?"
EV2H_LIB=exp/libev2h_TRACE.so timeout 300 python tools/fused_trace.py tf32x3 > gpurun_out/${TAG}_trace_f16.txt 2>&1; echo "trace rc=$?"
python - <<PY
import json
for n in ("f16","bf16"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_%s.json"%n).read().strip().splitlines()[-1])
        k=d["kernels"]
        print(n, "value %.0f ms %.3f e2e %.0f frac %.3f fused %.3f ms"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], k["ev2h_sa_msg_fused_tc"]["ms_per_step"]))
        print("   ", {a: round(b["ms_per_step"],3) for a,b in k.items()})
    except Exception as e: print(n, "failed", e)
PY
grep "====" gpurun_out/${TAG}_trace_f16.txt
