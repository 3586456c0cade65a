"""One encoder forward (B = 64, N = 2048) and one batch of event windows (B = 1024) for an ncu capture of the
non-MLP kernels: FPS, ball query (+ compaction), first-occurrence flags, window aggregation / sampling.
  ncu --set full --clock-control none -k regex:'fps_kernel|ball_query_kernel|first_occurrence|window_' -c 24 -o out python tools/ncu_small_kernels.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ev2hands_b200 as e2h  # noqa: E402
from ev2hands_b200 import synth  # noqa: E402
import bench  # noqa: E402

dev = torch.device("cuda:0")
enc = bench.build_encoder(dev)
ev = torch.from_numpy(synth.make_windows(64, 2048, seed=1236)).to(dev)
s1 = torch.from_numpy(synth.make_start_indices(64, 2048, 0)).to(dev)
s2 = torch.from_numpy(synth.make_start_indices(64, 512, 1)).to(dev)
with torch.no_grad():
    for _ in range(2):          # the second forward is the warm one
        enc(ev, fps_starts=(s1, s2))
torch.cuda.synchronize()
B, n = 1024, 2048
for mode, cols in (("stream", 0), ("erpc", 2)):
    raw = synth.make_raw_events(B * n // 2 + n, seed=1, duration=2.0e3 * (B // 2 + 1), extra_columns=cols)
    wb = e2h.EventWindowBuilder(mode)
    idx = torch.from_numpy(np.random.RandomState(0).randint(0, 1500, size=(B, 2048))).to(dev)
    d = torch.from_numpy(raw).to(dev)
    for _ in range(2):
        wb(d, np.arange(B) * (n // 2), np.full(B, n), sample_idx=idx)
    torch.cuda.synchronize()
print("done")
