"""FPS kernels for long windows: time of ev2h_fps_variant_f32 (1 exhaustive, 2 cluster, 3 pruned) per shape.  GPU only."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ev2hands_b200 import _capi, synth           # noqa: E402

dev = torch.device("cuda:0")
for n, s, mode in ((16384, 512, "events"), (16384, 512, "uniform"), (8192, 512, "events"), (6000, 512, "events")):
    base = synth.make_windows(16, n, seed=11, mode=mode)
    for b in (8, 64, 256):
        ev = torch.from_numpy(np.concatenate([base] * (b // 16 + 1))[:b]).to(dev)
        x = ev[:, :3, :]
        start = torch.from_numpy(synth.make_start_indices(b, n, seed=1)).to(dev)
        out = {}
        for variant in (1, 2, 3):
            if variant == 2 and b * (2 if n <= 8192 else 4) > 148 * 4:
                continue
            for _ in range(2):
                idx = _capi.fps(x, _capi.cf_strides(x), start, b, n, s, variant=variant)[0]
            torch.cuda.synchronize()
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                idx = _capi.fps(x, _capi.cf_strides(x), start, b, n, s, variant=variant)[0]
            z.record()
            torch.cuda.synchronize()
            out[variant] = (a.elapsed_time(z) / 3, idx)
        for _ in range(2):
            _capi.fps(x, _capi.cf_strides(x), start, b, n, 2, variant=3)
        torch.cuda.synchronize()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _capi.fps(x, _capi.cf_strides(x), start, b, n, 2, variant=3)
        z.record()
        torch.cuda.synchronize()
        setup = a.elapsed_time(z)
        same = all(torch.equal(out[1][1], v[1]) for v in out.values())
        print("N=%d S=%d %s B=%d: " % (n, s, mode, b) + "  ".join("v%d %.3f ms" % (k, v[0]) for k, v in out.items()) + "  (v3 with S=2: %.3f ms)  identical=%s" % (setup, same))
