"""How much the pruned FPS kernel skips: needs the stats build (tools/build_variant.sh fpsstats fps.cu -DEV2H_FPS_STATS,
EV2H_LIB=exp/libev2h_fpsstats.so).  Prints, per shape, the fraction of warp ranges entered and of buckets updated."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ev2hands_b200 import _capi, synth           # noqa: E402

dev = torch.device("cuda:0")
lib = _capi.lib()
lib.ev2h_debug_fps_stats.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_ulonglong * 4)()
for n, s, mode in ((16384, 512, "events"), (16384, 512, "uniform"), (16384, 128, "events"), (8192, 512, "events")):
    b = 8
    ev = torch.from_numpy(synth.make_windows(b, n, seed=11, mode=mode)).to(dev)
    x = ev[:, :3, :]
    start = torch.from_numpy(synth.make_start_indices(b, n, seed=1)).to(dev)
    lib.ev2h_debug_fps_stats(buf, 1)
    _capi.fps(x, _capi.cf_strides(x), start, b, n, s, variant=3)
    lib.ev2h_debug_fps_stats(buf, 1)
    nb = (32 if n > 8192 else 16) // 4
    it, entered, upd = buf[0], buf[1], buf[2]
    print("N=%d S=%d %s: warp iterations %d, ranges entered %.3f, buckets updated %.3f of all (%.2f per entered range)"
          % (n, s, mode, it, entered / it, upd / (it * nb), upd / max(entered, 1)))
