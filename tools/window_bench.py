"""Event-window construction on the device vs the reference's numpy recipe on the host.
Usage (GPU box): python tools/window_bench.py [--windows 1024] [--mode stream|erpc]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ev2hands_b200 as e2h  # noqa: E402
from ev2hands_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows", type=int, default=1024)
    ap.add_argument("--mode", default="stream")
    ap.add_argument("--events", type=int, default=2048, help="raw events per window")
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    B, n = args.windows, args.events
    ev = synth.make_raw_events(B * n // 2 + n, seed=1, duration=2.0e3 * (B // 2 + 1), extra_columns=2 if args.mode == "erpc" else 0)
    starts = np.arange(B) * (n // 2)                      # half-overlapping windows
    counts = np.full(B, n)
    idx = torch.from_numpy(np.random.RandomState(0).randint(0, 1500, size=(B, 2048)))
    dev = torch.device("cuda:0")
    d_ev, d_idx = torch.from_numpy(ev).to(dev), idx.to(dev)
    wb = e2h.EventWindowBuilder(args.mode)
    for _ in range(3):
        rec, n_pix, n_bad = wb.aggregate(d_ev, starts, counts)
        wb.sample(rec, n_pix, n_bad, d_idx)
    torch.cuda.synchronize()
    a, m, b = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t_agg = t_smp = 0.0
    for _ in range(args.steps):
        a.record()
        rec, n_pix, n_bad = wb.aggregate(d_ev, starts, counts)
        m.record()
        out = wb.sample(rec, n_pix, n_bad, d_idx)
        b.record()
        torch.cuda.synchronize()
        t_agg += a.elapsed_time(m)
        t_smp += m.elapsed_time(b)
    t_agg, t_smp = t_agg / args.steps, t_smp / args.steps
    # end to end with the reference's host-side draw (pixel counts read back, np.random.choice per window)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        wb(d_ev, starts, counts)
    torch.cuda.synchronize()
    t_e2e = (time.perf_counter() - t0) / 5 * 1e3
    # host: the reference's recipe (oracle/window_oracle.py: np.add.at grids, nonzero, draw, pc_normalize), a sample of the batch
    from oracle import window_oracle as wo
    nb = min(B, 32)
    t0 = time.perf_counter()
    for w in range(nb):
        r = wo.aggregate(ev[starts[w]:starts[w] + n], args.mode)
        wo.sample_normalize(r, np.random.choice(r.shape[0], 2048))
    t_cpu = (time.perf_counter() - t0) / nb * 1e3
    bytes_alg = B * (n * 32 + 1500 * 20 + 2048 * 20 * 2 + 2048 * 8)
    print(json.dumps({"metric": "event windows built/s", "mode": args.mode, "windows": B, "raw_events_per_window": n,
                      "aggregate_ms": t_agg, "sample_ms": t_smp, "windows_per_s_device": B / (t_agg + t_smp) * 1e3,
                      "windows_per_s_with_host_draw": B / t_e2e * 1e3, "cpu_oracle_ms_per_window": t_cpu,
                      "cpu_windows_per_s_1_thread": 1e3 / t_cpu,
                      "hbm": {"algorithmic_bytes": bytes_alg, "achieved_gbs": bytes_alg / (t_agg + t_smp) / 1e6},
                      "occupied_pixels_mean": float(n_pix.float().mean())}))


if __name__ == "__main__":
    main()
