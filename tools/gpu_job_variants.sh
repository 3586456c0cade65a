#!/bin/bash
# epilogue-set variants A/B on ONE box, interleaved twice: EV2H_FUSED_ES = 1 (one set), 2 (two sets in the one-CTA-per-SM instances)
TAG=${1:-j18}
mkdir -p gpurun_out
SUB="encoder or regressor or compact or sharding or long_window or batch64 or module_by_module"
EV2H_FUSED_ES=2 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -x -k "$SUB" > gpurun_out/${TAG}_pytest_es2.log 2>&1; echo "pytest es2 rc=$?"; tail -n 2 gpurun_out/${TAG}_pytest_es2.log
for rep in a b; do
  for v in 1 2; do
    EV2H_FUSED_ES=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline --no-raw-events > gpurun_out/${TAG}_bench_es$v$rep.json 2> gpurun_out/${TAG}_bench_es$v$rep.err; echo "bench es$v$rep rc=$?"
  done
done
python - <<PY
import json
for n in ("es1a","es2a","es1b","es2b"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_%s.json"%n).read().strip().splitlines()[-1])
        k=d["kernels"]
        print(n, "value %.0f ms %.3f e2e %.0f frac %.3f fused %.3f ms"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], k["ev2h_sa_msg_fused_tc"]["ms_per_step"]))
    except Exception as e: print(n, "failed", e)
PY
