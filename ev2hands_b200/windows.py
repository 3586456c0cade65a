"""Event windows on the GPU: raw camera events -> the encoder's [B, 5, N] input.

Host-side mirror of the window construction in the reference's dataset classes

    ``src/Ev2Hands/dataset/evaluation_stream.py:177-214``  ``ERPCParser.__getitem__``       (mode "stream")
    ``src/Ev2Hands/dataset/erpc.py:170-249``               ``Ev2HandSDataset.__getitem__``  (mode "erpc")

over ``ev2h_window_aggregate_f64`` / ``ev2h_window_sample_f32`` of ``libev2h.so``: per-pixel sums of a window's
events (what the reference does with ``np.add.at`` on 346x260 grids), the occupied pixels in the reference's order,
the draw of N of them with replacement and ``pc_normalize`` - for a whole batch of windows per call, the raw events
staying on the device.  The draw uses numpy's global generator exactly like the reference
(``np.random.choice(M, N)``, one call per window in batch order), so a seeded run selects the same points; that
needs the pixel counts M on the host, one small read-back per batch.  Pass ``sample_idx`` to skip it.

There is no CPU path: tensors must be CUDA tensors and a missing library raises.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _capi

SENSOR_W, SENSOR_H = 346, 260      # src/settings.py:21-22
_MODES = {"stream": _capi.WINDOW_STREAM, "erpc": _capi.WINDOW_ERPC}


class EventWindowBuilder:
    """``builder(events, starts, counts) -> float32 [B, 5, n_events]`` = (x, y, t, n_pos, n_neg) per point.

    ``events``  float64 ``[n, >= 4]`` CUDA tensor, rows (x, y, t, polarity, ...) in stream order - what
                ``get_events_by_time`` returns ("stream": t in ms, evaluation_stream.py:127-146) or rows of the
                HDF5 event table ("erpc": t in ns plus two bookkeeping columns, erpc.py:174-176);
    ``starts``, ``counts``  first row and number of rows of every window (windows may overlap, as the sliding
                ``dataset[index : index + 2048]`` of erpc.py:173-174 does).
    """

    def __init__(self, mode: str = "stream", n_events: int = 2048, width: int = SENSOR_W, height: int = SENSOR_H,
                 sampling: bool = True):
        """``sampling`` mirrors ``Ev2HandSDataset(sampling=...)`` (erpc.py:106, :216-226): True draws all ``n_events``
        points with replacement; False keeps every occupied pixel once and draws only the missing ``n_events - M``."""
        if mode not in _MODES:
            raise ValueError("mode must be 'stream' or 'erpc'")
        self.mode, self.n_events, self.width, self.height = mode, int(n_events), int(width), int(height)
        self.sampling = bool(sampling)
        self.last_n_pixels = self.last_n_bad = None

    def aggregate(self, events: torch.Tensor, starts, counts):
        """-> (records [B, max_count, 5], n_pixels int32 [B], n_bad int32 [B]) on the device."""
        starts_t = torch.as_tensor(starts, dtype=torch.int64)
        counts_t = torch.as_tensor(counts, dtype=torch.int32)
        if counts_t.is_cuda:
            raise RuntimeError("counts must be host values: the largest one sizes the record buffer")
        if starts_t.numel() == 0 or starts_t.shape != counts_t.shape:
            raise RuntimeError("starts and counts must be equally long, non-empty 1-d sequences")
        if int(counts_t.min()) < 1:
            raise RuntimeError("every window needs at least one event")
        if int((starts_t.cpu() + counts_t).max()) > events.shape[0] or int(starts_t.min()) < 0:
            raise IndexError("a window reaches outside the event table")
        return _capi.window_aggregate(events, starts_t, counts_t, int(counts_t.max()), self.width, self.height, _MODES[self.mode])

    def draw(self, n_pixels: torch.Tensor) -> torch.Tensor:
        """The reference's draw: ``np.random.choice(M_b, n_events)`` per window, in batch order, from numpy's
        global generator (evaluation_stream.py:210, erpc.py:217) -> int64 [B, n_events] (host, pinned)."""
        m = n_pixels.cpu().numpy()
        if (m < 1).any():
            raise RuntimeError("a window has no event inside the sensor")
        if self.sampling:
            idx = np.stack([np.random.choice(int(mb), self.n_events) for mb in m]).astype(np.int64)
        else:
            # erpc.py:219-226: the pixels themselves, then n_events - M of them again (a window with more occupied
            # pixels than n_events would keep them all there; a fixed-size batch cannot)
            if (m > self.n_events).any():
                raise RuntimeError("sampling=False needs at most n_events occupied pixels per window")
            idx = np.stack([np.concatenate([np.arange(int(mb)), np.random.choice(int(mb), self.n_events - int(mb))]) if mb < self.n_events
                            else np.arange(int(mb)) for mb in m]).astype(np.int64)
        return torch.from_numpy(idx).pin_memory()

    def sample(self, records, n_pixels, n_bad, sample_idx=None) -> torch.Tensor:
        if sample_idx is None:
            sample_idx = self.draw(n_pixels)
        sample_idx = torch.as_tensor(sample_idx)
        if sample_idx.dim() != 2 or sample_idx.shape[0] != records.shape[0]:
            raise RuntimeError("sample_idx must be [B, n_points]")
        return _capi.window_sample(records, n_pixels, sample_idx.to(records.device, non_blocking=True), self.width, self.height, n_bad)

    def __call__(self, events: torch.Tensor, starts, counts, sample_idx=None, check: bool = False) -> torch.Tensor:
        records, n_pixels, n_bad = self.aggregate(events, starts, counts)
        out = self.sample(records, n_pixels, n_bad, sample_idx)
        self.last_n_pixels, self.last_n_bad = n_pixels, n_bad
        if check and int(n_bad.sum()) != 0:       # numpy raises IndexError on such rows (or wraps negative ones)
            raise IndexError("%d events outside the %dx%d sensor or sample indices outside their window"
                             % (int(n_bad.sum()), self.width, self.height))
        return out
