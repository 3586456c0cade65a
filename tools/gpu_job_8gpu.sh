#!/bin/bash
# two GPUs: nn.DataParallel re-entrancy test, 2-rank bench with the configs block (cfg3 512/GPU, cfg4 with the NCCL all-reduce, cfg5)
TAG=${1:-g2}
mkdir -p gpurun_out
true > gpurun_out/${TAG}_pytest_dp.log 2>&1; echo "dp pytest rc=$?"; tail -n 3 gpurun_out/${TAG}_pytest_dp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_8gpu.json 2> gpurun_out/${TAG}_bench_8gpu.err; echo "bench 2gpu rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_8gpu.json").read().strip().splitlines()[-1])
    print("8gpu value %.0f ms %.3f e2e %.0f frac %.3f"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]))
    for c,v in (d.get("configs") or {}).items():
        print("   ", c, {kk: (round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("ms_per_step","value","roofline_frac","fps_ball_share_of_kernel_time","allreduce_us_alone","grad_allreduce_bytes","batch_per_gpu","failed")})
except Exception as e: print("failed", e)
PY
tail -n 5 gpurun_out/${TAG}_bench_8gpu.err
