"""Deterministic synthetic LiDAR clouds (tools/synth.cpp) for tests, smoke() and bench.py."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
_SRC = ROOT / "tools" / "synth.cpp"
_LIB_PATH = ROOT / "tools" / "libsynth.so"
_lib = None


def build() -> Path:
    if not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < _SRC.stat().st_mtime:
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", str(_LIB_PATH), str(_SRC), "-lpthread"],
                       check=True)
    return _LIB_PATH


def _L():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.synth_sensor_frame.restype = C.c_uint32
        _lib.synth_stress.restype = C.c_uint32
    return _lib


def make_frame(seed: int = 0, beams: int = 64, azimuth_steps: int = 2083, elev=(-24.8, 2.0), sensor_xy=(0.0, 0.0),
               n_boxes: int = 300, n_cyls: int = 200, scene_seed: int | None = None, threads: int | None = None):
    """(N,4) float32 x,y,z,intensity. Defaults = SURVEY config 5 (64-beam, ~120k returns)."""
    out = np.zeros((beams * azimuth_steps, 4), np.float32)
    threads = threads or min(8, os.cpu_count() or 1)
    n = _L().synth_sensor_frame(C.c_uint64(seed if scene_seed is None else scene_seed), C.c_uint64(seed),
                                C.c_uint32(beams), C.c_uint32(azimuth_steps), C.c_double(elev[0]), C.c_double(elev[1]),
                                C.c_double(sensor_xy[0]), C.c_double(sensor_xy[1]), C.c_uint32(n_boxes),
                                C.c_uint32(n_cyls), out.ctypes.data_as(C.POINTER(C.c_float)), C.c_uint32(threads))
    return out[:n].copy()


def make_frame_128(seed: int = 12345):
    """SURVEY config 3: 128 beams in [-25, +15] deg x 2048 azimuth steps (~260k returns)."""
    return make_frame(seed, beams=128, azimuth_steps=2048, elev=(-25.0, 15.0))


def make_stress(seed: int = 777, n_blobs: int = 1200, blob_pts: int = 250, sigma: float = 0.15, n_walls: int = 4,
                wall_len: float = 100.0, wall_height: float = 3.0, lattice: float = 0.05):
    cap = n_blobs * blob_pts + n_walls * int(wall_len / lattice) * int(wall_height / lattice)
    out = np.zeros((max(cap, 1), 4), np.float32)
    n = _L().synth_stress(C.c_uint64(seed), C.c_uint32(n_blobs), C.c_uint32(blob_pts), C.c_double(sigma),
                          C.c_uint32(n_walls), C.c_double(wall_len), C.c_double(wall_height), C.c_double(lattice),
                          out.ctypes.data_as(C.POINTER(C.c_float)))
    return out[:n].copy()


def make_merged_1m(seed: int = 777):
    """SURVEY config 4 (scaled to ~1M points): four sensors sharing one scene + dense stress structures."""
    parts = [make_frame(seed + i, beams=64, azimuth_steps=1024, sensor_xy=xy, scene_seed=seed)
             for i, xy in enumerate([(1.0, 0.5), (1.0, -0.5), (-1.0, 0.5), (-1.0, -0.5)])]
    parts.append(make_stress(seed))
    return np.concatenate(parts, axis=0)
