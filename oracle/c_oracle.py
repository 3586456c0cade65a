"""TEST INFRASTRUCTURE - ctypes loader for the plain-C oracle (sa_oracle_c.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this.  ``build()`` compiles oracle/_build/liboracle.so with the Makefile here.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "sa_oracle_c.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def fps(xyz: np.ndarray, n_sample: int, start: np.ndarray) -> np.ndarray:
    """xyz [B,N,3] f32, start [B] -> int64 [B,n_sample]."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    B, N, _ = xyz.shape
    out = np.empty((B, n_sample), dtype=np.int64)
    for b in range(B):
        lib().orc_fps(_fp(xyz[b]), ctypes.c_int64(N), ctypes.c_int64(n_sample),
                      ctypes.c_int64(int(start[b])), _ip(out[b]))
    return out


def sqdist(centres: np.ndarray, xyz: np.ndarray) -> np.ndarray:
    """centres [B,S,3], xyz [B,N,3] -> f32 [B,S,N]."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    centres = np.ascontiguousarray(centres, dtype=np.float32)
    B, N, _ = xyz.shape
    S = centres.shape[1]
    out = np.empty((B, S, N), dtype=np.float32)
    for b in range(B):
        lib().orc_sqdist(_fp(centres[b]), ctypes.c_int64(S), _fp(xyz[b]), ctypes.c_int64(N), _fp(out[b]))
    return out


def ball_query(radius: float, n_neighbor: int, xyz: np.ndarray, centres: np.ndarray) -> np.ndarray:
    """xyz [B,N,3], centres [B,S,3] -> int64 [B,S,n_neighbor]."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    centres = np.ascontiguousarray(centres, dtype=np.float32)
    B, N, _ = xyz.shape
    S = centres.shape[1]
    out = np.empty((B, S, n_neighbor), dtype=np.int64)
    r2 = ctypes.c_float(float(np.float32(radius ** 2)))
    for b in range(B):
        lib().orc_ball_query(_fp(xyz[b]), ctypes.c_int64(N), _fp(centres[b]), ctypes.c_int64(S),
                             r2, ctypes.c_int64(n_neighbor), _ip(out[b]))
    return out


def mlp_max_f64(x: np.ndarray, layers, eps: float = 1e-5) -> np.ndarray:
    """x [G,K,C0] f32; layers = [(W[Co,Ci], b, gamma, beta, mean, var), ...] ->
    float64 [G, C_last], evaluated entirely in double precision."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    G, K, C0 = x.shape
    L = len(layers)
    dims = np.array([C0] + [int(l[0].shape[0]) for l in layers], dtype=np.int64)
    keep = [[np.ascontiguousarray(np.asarray(t, dtype=np.float32).reshape(t.shape[0], -1) if i == 0
                                  else np.asarray(t, dtype=np.float32)) for i, t in enumerate(l)]
            for l in layers]
    PF = ctypes.POINTER(ctypes.c_float)
    cols = []
    for i in range(6):
        arr = (PF * L)(*[_fp(keep[l][i]) for l in range(L)])
        cols.append(arr)
    out = np.empty((G, int(dims[-1])), dtype=np.float64)
    lib().orc_mlp_max_f64(_fp(x), ctypes.c_int64(G), ctypes.c_int64(K), ctypes.c_int(L), _ip(dims),
                          *cols, ctypes.c_double(eps),
                          out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return out
