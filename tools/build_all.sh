#!/bin/bash
# builds libev2h.so and the EV2H_FUSED_TRACE variant (exp/libev2h_TRACE.so, used by tools/fused_trace.py)
set -e
cd "$(dirname "$0")/.."
python -c "from ev2hands_b200 import build as b; print(b.build(force=True))" 2>&1 | tail -1
python - <<'PY'
import glob, os, subprocess
srcs=sorted(glob.glob('ev2hands_b200/csrc/*.cu'))
flags=["-gencode","arch=compute_100a,code=sm_100a","-lineinfo","-O3","-std=c++17","-Xcompiler","-fPIC","-Xcompiler","-fvisibility=hidden","-DEV2H_FUSED_TRACE"]
os.makedirs('exp/build_trace',exist_ok=True)
procs=[]
for s in srcs:
    o='exp/build_trace/'+os.path.basename(s)[:-3]+'.o'
    procs.append((o,subprocess.Popen(['nvcc']+flags+['-c',s,'-o',o],stderr=subprocess.DEVNULL)))
objs=[]
for o,p in procs:
    assert p.wait()==0, o; objs.append(o)
subprocess.check_call(['nvcc','-shared','-gencode','arch=compute_100a,code=sm_100a','-o','exp/libev2h_TRACE.so']+objs+['-lcudart'])
print('trace lib ok')
PY
