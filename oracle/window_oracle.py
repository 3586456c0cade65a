"""CPU restatement (numpy / torch) of the reference's event-window construction - TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import this module; the product
path (``ev2hands_b200.windows`` -> ``libev2h.so``) never does.

Two builders exist in the reference:

  "stream"  ``src/Ev2Hands/dataset/evaluation_stream.py:177-214`` (``ERPCParser.__getitem__``): raw rows
            (x, y, t [ms], p); t is rebased to the first event (:188), events are summed per pixel with
            ``np.add.at`` into float32 grids (:190-200), the occupied pixels are listed in ``np.nonzero`` order
            (:203-208), N of them are drawn with replacement (:210-211) and x / y / t are normalised (:215,
            ``pc_normalize`` :12-29).
  "erpc"    ``src/Ev2Hands/dataset/erpc.py:170-249`` (``Ev2HandSDataset.__getitem__``): the same per-pixel sums
            (:178-196), mean time scaled by 1e-6 (:192), pixels sorted by mean time and rebased to the earliest
            (:210-214), the same draw (:216-218) and normalisation (:249, ``pc_normalize`` :23-39).

Pinned against the reference itself: ``tests/golden/make_window_golden.py`` drives both reference methods
unmodified (missing third-party imports stubbed) and ``tests/test_oracle_cpu.py`` compares this module with the
saved outputs bit for bit.  The one place the reference leaves open is the order of pixels with EQUAL mean time in
"erpc" (``np.argsort`` default, an unstable sort, erpc.py:210); this restatement and the CUDA path keep pixel order
among equals (stable), which is one of the orders the reference may produce.
"""
from __future__ import annotations

import numpy as np

WIDTH, HEIGHT = 346, 260          # src/settings.py:21-22


def aggregate(events: np.ndarray, mode: str, width: int = WIDTH, height: int = HEIGHT) -> np.ndarray:
    """raw rows [n, >=4] float64 -> per-pixel records float32 [M, 5] = (x, y, t_mean, n_pos, n_neg) in the order
    the reference holds them right before the random draw."""
    ev = np.asarray(events, dtype=np.float64)
    x = ev[:, 0].astype(np.int32)
    y = ev[:, 1].astype(np.int32)
    t = ev[:, 2].copy()
    p = ev[:, 3]
    if mode == "stream":
        t -= t[0]                                           # evaluation_stream.py:188
    sums = np.zeros((height, width, 3), dtype=np.float32)
    cnt = np.zeros((height, width), dtype=np.float32)
    # float32 grids fed with float64 values: every single addition is done in double and rounded to float32
    np.add.at(sums, (y, x, 0), t)
    np.add.at(sums, (y, x, 1), p == 1)
    np.add.at(sums, (y, x, 2), p != 1)
    np.add.at(cnt, (y, x), 1)
    yi, xi = np.nonzero(cnt)                                # row-major pixel order
    t_mean = sums[yi, xi, 0] / cnt[yi, xi]
    if mode == "erpc":
        t_mean = t_mean * 1e-6                              # erpc.py:192 (float32 x weak python scalar)
    rec = np.stack([xi.astype(np.float32), yi.astype(np.float32), t_mean.astype(np.float32),
                    sums[yi, xi, 1], sums[yi, xi, 2]], axis=1).astype(np.float32)
    if mode == "erpc":
        rec = rec[np.argsort(rec[:, 2], kind="stable")]    # erpc.py:210-211 (see the module docstring on ties)
        rec[:, 2] -= rec[0, 2]                              # erpc.py:214
    elif mode != "stream":
        raise ValueError("mode must be 'stream' or 'erpc'")
    return rec


def sample_normalize(records: np.ndarray, sample_idx: np.ndarray, width: int = WIDTH, height: int = HEIGHT) -> np.ndarray:
    """records [M,5], indices [N] (the reference's ``np.random.choice(M, N)``) -> window float32 [5, N]."""
    import torch
    ev = torch.tensor(records[np.asarray(sample_idx)], dtype=torch.float32)
    pc = ev[:, :3]
    pc[:, 0] /= width
    pc[:, 1] /= height
    pc[:, :2] = 2 * pc[:, :2] - 1
    ts = pc[:, 2:]
    t_max, t_min = ts.max(0).values, ts.min(0).values
    pc[:, 2:] = (2 * ((ts - t_min) / (t_max - t_min))) - 1
    return ev.permute(1, 0).contiguous().numpy()


def build_windows(events: np.ndarray, starts, counts, sample_idx: np.ndarray, mode: str) -> np.ndarray:
    """B windows -> float32 [B, 5, N]; sample_idx int64 [B, N]."""
    out = [sample_normalize(aggregate(events[s:s + c], mode), sample_idx[b]) for b, (s, c) in enumerate(zip(starts, counts))]
    return np.stack(out)
