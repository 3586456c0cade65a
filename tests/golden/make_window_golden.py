"""Generate tests/golden/windows.npz from the REAL reference's event-window builders.

Run in the build container only (reads /root/reference):

    python tests/golden/make_window_golden.py

``dataset/evaluation_stream.py`` and ``dataset/erpc.py`` import packages that are absent here (``dv``, ``h5py``)
and ``settings`` (which imports pyrender); those imports are satisfied with empty stand-in modules carrying the
three constants the builders read (src/settings.py:21-23).  The two ``__getitem__`` methods then run UNMODIFIED on
objects made with ``object.__new__`` whose data attributes are filled with synthetic raw events
(``ev2hands_b200.synth.make_raw_events``).  ``np.random.choice`` is wrapped only to record the indices it returns.
Nothing of the reference is copied; only inputs and outputs are saved.
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from ev2hands_b200 import synth  # noqa: E402

REF_DIR = "/root/reference/src/Ev2Hands/dataset"


def load_reference():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
    stub("settings", OUTPUT_WIDTH=346, OUTPUT_HEIGHT=260, LNES_WINDOW_MS=5)     # src/settings.py:21-23
    stub("dv", AedatFile=None)
    stub("h5py")
    stub("camera", undistort=None, opencv_camera_view_to_screen_space_transform=None)
    pkg = types.ModuleType("refdataset")
    pkg.__path__ = [REF_DIR]
    sys.modules["refdataset"] = pkg
    mods = {}
    for name in ("augmentations", "evaluation_stream", "erpc"):
        spec = importlib.util.spec_from_file_location("refdataset." + name, os.path.join(REF_DIR, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules["refdataset." + name] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return mods["evaluation_stream"], mods["erpc"]


class LegacyIndexArray(np.ndarray):
    """``evaluation_stream.py:207`` indexes with a LIST holding Ellipsis and None (``n_evn[[..., None]]``), which
    NumPy < 1.23 (the reference pins an old NumPy, ev2hands.yml) read as the tuple ``[..., None]`` and NumPy 2
    rejects.  Arrays made by ``np.zeros`` inside that module are of this subclass, which restores the old reading."""

    def __getitem__(self, key):
        if isinstance(key, list) and any(k is Ellipsis or k is None for k in key):
            key = tuple(key)
        return super().__getitem__(key)


class LegacyNumpy:
    """stands in for the name ``np`` inside the reference module: numpy itself, except that zeros() returns LegacyIndexArray."""

    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def zeros(*a, **kw):
        return np.zeros(*a, **kw).view(LegacyIndexArray)


class ChoiceRecorder:
    def __init__(self):
        self.calls = []
        self.orig = np.random.choice

    def __call__(self, a, size=None, *args, **kw):
        r = self.orig(a, size, *args, **kw)
        self.calls.append((int(a), np.asarray(r).copy()))
        return r


def main():
    stream_mod, erpc_mod = load_reference()
    stream_mod.np = LegacyNumpy()
    rec = ChoiceRecorder()
    np.random.choice = rec
    out = {}

    # ---- "stream": ERPCParser.__getitem__ (evaluation_stream.py:177-214) over a microsecond stream ---------------
    raw_us = synth.make_raw_events(9000, seed=11, t0=1.7e9, duration=9000.0)     # ~1 event / us: windows of > 2 ms
    parser = object.__new__(stream_mod.ERPCParser)
    parser.events = raw_us.copy()
    parser.joints = np.zeros([1, 2, 21, 3])
    parser.camera = {}
    parser.e_id = 0
    parser.n_events = 0
    np.random.seed(5)
    starts, counts, wins = [], [], []
    for _ in range(3):
        e0 = parser.e_id
        item = parser[0]
        n_used = rec.calls[-1]                  # (M, indices) of this window
        # rows the window was built from: the method read events until > 2 ms and >= 2048 rows (:127-146)
        wins.append(item["data"].numpy())
        starts.append(e0)
        counts.append(None)
        out.setdefault("stream_M", []).append(n_used[0])
        out.setdefault("stream_idx", []).append(n_used[1])
    # recover each window's row count by replaying the reader's stopping rule on the same stream
    ts_ms = raw_us[:, 2] * 1e-3
    for i, s in enumerate(starts):
        n = 1
        while True:
            if abs(ts_ms[s + n] - ts_ms[s]) > stream_mod.WINDOWS_SIZE and n >= 2048:
                break
            n += 1
        counts[i] = n
    ev_ms = raw_us[:, :4].copy()
    ev_ms[:, 2] = raw_us[:, 2] * 1e-3           # what get_event hands out (:104)
    out.update(stream_events=ev_ms, stream_starts=np.array(starts), stream_counts=np.array(counts),
               stream_windows=np.stack(wins), stream_M=np.array(out["stream_M"]), stream_idx=np.stack(out["stream_idx"]))

    # ---- "erpc": Ev2HandSDataset.__getitem__ (erpc.py:170-249), sampling on, augmentation off ---------------------
    # "erpc":  small, distinct timestamps - every pixel's mean time is distinct in float32, so the reference's
    #          unstable argsort (:210) has one possible result;
    # "erpct": nanosecond timestamps around 3e9 - float32 means collide, and the order among equal means is the
    #          sort's choice (tests compare what does not depend on it).
    hand = {'global_orient': np.zeros(3), 'hand_pose': np.zeros(45), 'shape': np.zeros(10), 'trans': np.zeros(3)}
    for tag, kw, seed in (("erpc", dict(t0=0.0, duration=1.2e7, unique_times=True), 6),
                          ("erpct", dict(t0=3.0e9, duration=1.5e7), 7)):
        rows = synth.make_raw_events(6000, seed=12, extra_columns=2, **kw)                        # t in ns
        rows[:, 5] = np.random.RandomState(1).randint(0, 4, size=rows.shape[0])                   # event labels
        ds = object.__new__(erpc_mod.Ev2HandSDataset)
        ds.dataset = rows
        ds.annotations = {0: {'left': dict(hand), 'right': dict(hand)}}
        ds.augment = False
        ds.sampling = True
        ds.demo = False
        np.random.seed(seed)
        e_starts, e_wins, e_M, e_idx = [0, 1500, 3952], [], [], []
        for s in e_starts:
            n0 = len(rec.calls)
            item = ds[s]
            assert len(rec.calls) == n0 + 1
            e_wins.append(item["events"].numpy())
            e_M.append(rec.calls[-1][0])
            e_idx.append(rec.calls[-1][1])
        out.update({tag + "_events": rows, tag + "_starts": np.array(e_starts), tag + "_counts": np.array([2048] * 3),
                    tag + "_windows": np.stack(e_wins), tag + "_M": np.array(e_M), tag + "_idx": np.stack(e_idx)})

    # ---- "erpcpad": the same class with sampling=False (erpc.py:219-226): every occupied pixel once, in time order,
    # then n_events - M more drawn with replacement ------------------------------------------------------------------
    rows = synth.make_raw_events(6000, seed=13, extra_columns=2, t0=0.0, duration=1.2e7, unique_times=True)
    rows[:, 5] = np.random.RandomState(2).randint(0, 4, size=rows.shape[0])
    ds = object.__new__(erpc_mod.Ev2HandSDataset)
    ds.dataset = rows
    ds.annotations = {0: {'left': dict(hand), 'right': dict(hand)}}
    ds.augment = False
    ds.sampling = False
    ds.demo = False
    np.random.seed(8)
    p_starts, p_wins, p_M, p_idx = [10, 2500], [], [], []
    for s in p_starts:
        n0 = len(rec.calls)
        item = ds[s]
        assert len(rec.calls) == n0 + 1
        m, extra = rec.calls[-1]
        p_wins.append(item["events"].numpy())
        p_M.append(m)
        p_idx.append(np.concatenate([np.arange(m), extra]))
    out.update({"erpcpad_events": rows, "erpcpad_starts": np.array(p_starts), "erpcpad_counts": np.array([2048] * 2),
                "erpcpad_windows": np.stack(p_wins), "erpcpad_M": np.array(p_M), "erpcpad_idx": np.stack(p_idx)})

    np.random.choice = rec.orig
    path = os.path.join(HERE, "windows.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
