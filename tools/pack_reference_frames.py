#!/usr/bin/env python
"""Pack the reference's 154 data/*.pcd frames into one BIT-lossless, compact cache file.

The frames are mm-quantised (float32(round(v*1000)/1000) reproduces every coordinate VALUE;
intensity is k/100), so each frame is stored as int32 millimetre deltas in the ORIGINAL point order
plus a uint8 intensity, followed by an exception list {flat word index, raw float32 bits} for every
word whose BIT pattern the millimetre form does not reproduce (the -0.0 coordinates: 12 words in
frame 0), LZMA-compressed (~60 MB for all frames). The round trip is asserted on the uint32 view.

The cache (data_cache/frames_mm.xz) is git-ignored but NOT gpurun-ignored, so bench.py can run the
reference's own sequence on the GPU box where /root/reference does not exist. Run here:
    python tools/pack_reference_frames.py
"""
import lzma
import os
import struct
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import oracle as O  # noqa: E402  (PCD reader only)

MAGIC_V1 = b"LB2FRM01"
MAGIC_V2 = b"LB2FRM02"  # v2: one xz stream per frame (value-lossless only: -0.0 came back as +0.0)
MAGIC = b"LB2FRM03"  # v3: v2 + per-frame exception list, bit-lossless


def _encode(p):
    pts = O.read_pcd(p)
    q = np.round(pts[:, :3].astype(np.float64) * 1000.0).astype(np.int32)
    back = (q.astype(np.float64) / 1000.0).astype(np.float32)
    assert np.array_equal(back, pts[:, :3]), f"{p}: not mm-exact"
    qi = np.round(pts[:, 3].astype(np.float64) * 100.0).astype(np.int32)
    assert np.array_equal((qi / 100.0).astype(np.float32), pts[:, 3]) and qi.min() >= 0 and qi.max() < 256
    dq = np.diff(q, axis=0, prepend=np.zeros((1, 3), np.int32)).T.copy()
    body = dq.tobytes() + qi.astype(np.uint8).tobytes()
    # words whose bit pattern differs from what the millimetre form decodes to (signed zeros)
    want = np.ascontiguousarray(pts, np.float32).view(np.uint32).reshape(-1)
    got = _decode_raw(pts.shape[0], body).view(np.uint32).reshape(-1)
    exc = np.nonzero(want != got)[0].astype(np.uint32)
    body += np.uint32(exc.size).tobytes() + exc.tobytes() + want[exc].tobytes()
    assert np.array_equal(_decode_raw(pts.shape[0], body).view(np.uint32).reshape(-1), want), f"{p}: not bit-exact"
    return pts.shape[0], lzma.compress(body, preset=6)


def pack(paths, out_path):
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:  # lzma releases the GIL
        blobs = list(ex.map(_encode, paths))
    with open(out_path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(blobs)))
        f.write(np.array([n for n, _ in blobs], np.uint32).tobytes())
        f.write(np.array([len(b) for _, b in blobs], np.uint32).tobytes())
        for _, b in blobs:
            f.write(b)
    return sum(len(b) for _, b in blobs)


def _decode_raw(n, raw):
    dq = np.frombuffer(raw, np.int32, 3 * n, 0).reshape(3, n)
    qi = np.frombuffer(raw, np.uint8, n, 12 * n)
    q = np.cumsum(dq, axis=1, dtype=np.int64).T
    out = np.empty((n, 4), np.float32)
    out[:, :3] = (q.astype(np.float64) / 1000.0).astype(np.float32)
    out[:, 3] = (qi.astype(np.float64) / 100.0).astype(np.float32)
    if len(raw) > 13 * n:  # v3 exception list
        (n_exc,) = struct.unpack_from("<I", raw, 13 * n)
        idx = np.frombuffer(raw, np.uint32, n_exc, 13 * n + 4)
        bits = np.frombuffer(raw, np.uint32, n_exc, 13 * n + 4 + 4 * n_exc)
        out.view(np.uint32).reshape(-1)[idx] = bits
    return out


def unpack(path):
    """Returns a list of (N,4) float32 arrays, bit-identical (as uint32 words) to the PCD files for v3 containers.
    Also reads the older value-lossless containers: v1 = one xz stream over all frames, v2 = one stream per frame."""
    with open(path, "rb") as f:
        magic = f.read(8)
        assert magic in (MAGIC, MAGIC_V2, MAGIC_V1), f"{path}: not a frame cache"
        (nf,) = struct.unpack("<I", f.read(4))
        counts = [int(n) for n in np.frombuffer(f.read(4 * nf), np.uint32)]
        if magic in (MAGIC, MAGIC_V2):
            sizes = np.frombuffer(f.read(4 * nf), np.uint32)
            blobs = [f.read(int(sz)) for sz in sizes]
        else:
            raw, blobs, pos = lzma.decompress(f.read()), [], 0
            for n in counts:
                blobs.append(raw[pos:pos + 13 * n])
                pos += 13 * n
    if magic == MAGIC_V1:
        return [_decode_raw(n, b) for n, b in zip(counts, blobs)]
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
        return list(ex.map(lambda a: _decode_raw(a[0], lzma.decompress(a[1])), zip(counts, blobs)))


if __name__ == "__main__":
    paths = O.reference_frame_paths()
    if not paths:
        sys.exit("reference data not found")
    out = ROOT / "data_cache" / "frames_mm.xz"
    out.parent.mkdir(exist_ok=True)
    size = pack(paths, out)
    print(f"packed {len(paths)} frames -> {out} ({size/1e6:.1f} MB)")
    frames = unpack(out)
    for p, fr in zip(paths[:3] + paths[-2:], frames[:3] + frames[-2:]):
        assert np.array_equal(O.read_pcd(p).view(np.uint32), fr.view(np.uint32))
    print("round-trip verified (uint32 words)")
