#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -q > gpurun_out/j2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j2_pytest.log
tail -5 gpurun_out/j2_pytest.log
EV2H_LIB=exp/libev2h_TRACE.so timeout 300 python tools/fused_trace.py tf32x3 > gpurun_out/j2_trace_f16.txt 2>&1; echo "trace rc=$?"
grep -c . gpurun_out/j2_trace_f16.txt
