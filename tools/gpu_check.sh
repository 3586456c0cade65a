#!/bin/bash
# Run on the GPU box via gpurun: parity tests, smoke, bench, ncu launch list.  Output in gpurun_out/<tag>_*.
TAG=${1:-run}
mkdir -p gpurun_out
(timeout 300 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | tail -40) > gpurun_out/${TAG}_pytest.log
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5) > gpurun_out/${TAG}_smoke.log
for P in ${MLPS:-fp32 tf32x3 bf16}; do
(timeout 150 python bench.py --steps 5 --warmup 3 --mlp $P 2>>gpurun_out/${TAG}_bench.err | tail -1) > gpurun_out/${TAG}_bench_$P.json
done
if [ "$4" == "default" ]; then
  # the driver's invocation (CUDA graph replay, CPU + stock-PyTorch-on-GPU baselines) and the decoder secondary number
  (timeout 300 python bench.py 2>>gpurun_out/${TAG}_bench.err | tail -1) > gpurun_out/${TAG}_bench_default.json
  (timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --with-decoder 2>>gpurun_out/${TAG}_bench.err | tail -1) > gpurun_out/${TAG}_bench_decoder.json
  (timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-graph 2>>gpurun_out/${TAG}_bench.err | tail -1) > gpurun_out/${TAG}_bench_eager.json
fi
if [ "$2" == "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --mlp ${NCU_MLP:-tf32x3} > gpurun_out/${TAG}_ncu_bench.log 2>&1
fi
tail -3 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_smoke.log; cat gpurun_out/${TAG}_bench_*.json | cut -c1-1500; tail -3 gpurun_out/${TAG}_bench.err
if [ "$3" == "full" ]; then
  # one --set full capture of the five fused launches of a step (skip the warm-up steps' launches)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sa_fused -s 15 -c 5 -o gpurun_out/${TAG}_fused -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --mlp ${NCU_MLP:-tf32x3} > gpurun_out/${TAG}_ncu_full.log 2>&1
  ncu -i gpurun_out/${TAG}_fused.ncu-rep --page raw --csv > gpurun_out/${TAG}_fused_raw.csv 2>/dev/null
fi
