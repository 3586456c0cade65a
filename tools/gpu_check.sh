#!/bin/bash
# Run on the GPU box via gpurun: parity tests, smoke, bench, ncu launch list.  Output in gpurun_out/<tag>_*.
TAG=${1:-run}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | tail -40) > gpurun_out/${TAG}_pytest.log
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5) > gpurun_out/${TAG}_smoke.log
(timeout 400 python bench.py --steps 5 --warmup 3 2>gpurun_out/${TAG}_bench.err | tail -1) > gpurun_out/${TAG}_bench.json
if [ "$2" == "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
fi
tail -3 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_smoke.log; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
