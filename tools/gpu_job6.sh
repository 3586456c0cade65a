#!/bin/bash
# full GPU suite + headline / configs bench + pool-placement variant + timeline
TAG=${1:-j9}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-raw-events > gpurun_out/${TAG}_bench_def.json 2> gpurun_out/${TAG}_bench_def.err; echo "bench def rc=$?"
EV2H_POOL_LOADER=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline --no-raw-events > gpurun_out/${TAG}_bench_pl.json 2> gpurun_out/${TAG}_bench_pl.err; echo "bench pool-loader rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline --no-raw-events --mlp bf16 > gpurun_out/${TAG}_bench_bf16.json 2> gpurun_out/${TAG}_bench_bf16.err
EV2H_LIB=exp/libev2h_TRACE.so timeout 300 python tools/fused_trace.py tf32x3 > gpurun_out/${TAG}_trace.txt 2>&1; echo "trace rc=$?"
python - <<PY
import json
for n in ("def","pl","bf16"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_%s.json"%n).read().strip().splitlines()[-1])
        k=d["kernels"]
        print(n, "value %.0f ms %.3f e2e %.0f frac %.3f fused %.3f ms"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], k["ev2h_sa_msg_fused_tc"]["ms_per_step"]))
        print("   ", {a: round(b["ms_per_step"],3) for a,b in k.items()})
        for c,v in (d.get("configs") or {}).items():
            print("   ", c, {kk: (round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("ms_per_step","value","roofline_frac","fps_ball_share_of_kernel_time","allreduce_us_alone","failed")})
            if "kernels_ms_per_step" in v: print("        ", {a: round(b,3) for a,b in v["kernels_ms_per_step"].items()})
        if "sustained" in d: print("   sustained", d["sustained"]["value"], d["sustained"]["roofline_frac"], d["sustained"]["clocks"])
    except Exception as e: print(n, "failed", e)
PY
grep "====" gpurun_out/${TAG}_trace.txt
