#!/bin/bash
# round-2 GPU job 1: full GPU test suite, bench of both fp32-level splits, launch list + ncu of the fused kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/j1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/j1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j1_pytest.log
tail -5 gpurun_out/j1_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/j1_bench_f16.json 2> gpurun_out/j1_bench_f16.err; echo "bench f16 rc=$?"
EV2H_SPLIT=tf32 timeout 300 python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline --no-raw-events > gpurun_out/j1_bench_tf32.json 2> gpurun_out/j1_bench_tf32.err; echo "bench tf32 rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline --no-raw-events --mlp bf16 > gpurun_out/j1_bench_bf16.json 2> gpurun_out/j1_bench_bf16.err; echo "bench bf16 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sa_fused -s 15 -c 5 -o gpurun_out/j1_fused python bench.py --steps 1 --warmup 3 --no-graph --no-configs --no-cpu-baseline --no-raw-events > gpurun_out/j1_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for n in ("f16","tf32","bf16"):
    try:
        d=json.loads(open("gpurun_out/j1_bench_%s.json"%n).read().strip().splitlines()[-1])
        k=d["kernels"]
        print(n, "value %.0f ms %.3f e2e %.0f frac %.3f fused %.3f ms"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], k["ev2h_sa_msg_fused_tc"]["ms_per_step"]))
        print("   ", {a: round(b["ms_per_step"],3) for a,b in k.items()})
        if "configs" in d: print(json.dumps(d["configs"])[:1500]); print(json.dumps(d.get("sustained")))
    except Exception as e: print(n, "failed", e)
PY
