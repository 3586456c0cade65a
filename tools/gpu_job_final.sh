#!/bin/bash
# final record of the round: full GPU suite, default bench line (with cpu baseline + configs), bf16, launch list, ncu of the fused kernel, timeline
TAG=${1:-fin}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; echo "bench default rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "bench reference rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline --no-raw-events --mlp bf16 > gpurun_out/${TAG}_bench_bf16.json 2> gpurun_out/${TAG}_bench_bf16.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline --no-raw-events --with-decoder > gpurun_out/${TAG}_bench_decoder.json 2> gpurun_out/${TAG}_bench_decoder.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-configs --no-cpu-baseline --no-raw-events > gpurun_out/${TAG}_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sa_fused -s 15 -c 5 -o gpurun_out/${TAG}_fused python bench.py --steps 1 --warmup 3 --no-graph --no-configs --no-cpu-baseline --no-raw-events > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu fused rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:'fps_kernel|ball_query|linear_tc|first_occ|attn_|three_' -c 12 -o gpurun_out/${TAG}_small python bench.py --steps 1 --warmup 1 --no-graph --no-configs --no-cpu-baseline --no-raw-events > gpurun_out/${TAG}_ncu_small.log 2>&1; echo "ncu small rc=$?"
EV2H_LIB=exp/libev2h_TRACE.so timeout 300 python tools/fused_trace.py tf32x3 > gpurun_out/${TAG}_trace.txt 2>&1; echo "trace rc=$?"
python - <<PY
import json
for n in ("default","bf16","reference","decoder"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "value %.1f ms %.3f"%(d["value"], d["ms_per_step"]), "e2e", d.get("e2e",{}).get("value"), "frac", (d.get("roofline") or {}).get("frac"))
        for c,v in (d.get("configs") or {}).items():
            print("   ", c, {kk: (round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("ms_per_step","value","roofline_frac","fps_ball_share_of_kernel_time","failed")})
        if "sustained" in d: print("   sustained", d["sustained"]["value"], d["sustained"]["roofline_frac"])
        if "cpu_baseline" in d: print("   cpu", json.dumps(d["cpu_baseline"])[:500])
        if "secondary" in d: print("   secondary", d["secondary"])
        if "from_raw_events" in d: print("   raw", json.dumps(d["from_raw_events"])[:400])
    except Exception as e: print(n, "failed", e)
PY
