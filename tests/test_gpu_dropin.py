"""The C++ drop-in classes (lidar-processing_b200/dropin) driven the way Processor::process drives
the reference's (reference src/processor.cpp:150-195), compared with the oracle."""
import subprocess

import numpy as np
import pytest

import oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_cpp_dropin_classes(pkg, ctx, golden_frames, tmp_path):
    import __graft_entry__ as ge

    exe = ge.build_dropin_test()
    pts = golden_frames[0]
    inp = tmp_path / "points.f32"
    pts.astype(np.float32).tofile(inp)
    prefix = tmp_path / "out"
    r = subprocess.run([str(exe), str(inp), str(prefix)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    seg = np.fromfile(f"{prefix}.seg.u32", np.uint32)
    ground = np.fromfile(f"{prefix}.ground.f32", np.float32).reshape(-1, 4)
    obstacle = np.fromfile(f"{prefix}.obstacle.f32", np.float32).reshape(-1, 4)
    clusters = np.fromfile(f"{prefix}.clusters.i32", np.int32)
    # same answer as the Python mirror over the same C ABI
    labels, gi, oi = ctx.segment(pts)
    assert np.array_equal(seg, labels)
    assert np.array_equal(ground.view(np.uint32), pts[gi].view(np.uint32))
    assert np.array_equal(obstacle.view(np.uint32), pts[oi].view(np.uint32))
    H.check_segmentation(pts, seg, gi, oi)
    H.check_clustering(obstacle, clusters)
    # outlines of the split clusters (Clusterer::outline_last_clusters) against the reference's host functions
    parts = [c for c, _ in O.split_clusters(obstacle, clusters)]
    for mode, tag in ((0, "convex"), (1, "concave_small"), (2, "concave")):
        sizes = np.fromfile(f"{prefix}.{tag}.sizes.u32", np.uint32)
        xy = np.fromfile(f"{prefix}.{tag}.xy.f32", np.float32).reshape(-1, 2)
        host_ids = np.fromfile(f"{prefix}.{tag}.host_ids.u32", np.uint32)
        assert sizes.size == len(parts)
        if mode == 2:  # findOrderedConcaveOutlines as a whole: the unmodified reference, else the host-compiled core
            want = O.ref_outlines(parts, 1) if O.ref_hull_available() else None
            if want is None:
                small = [w for w, _ in O.convex_outlines(parts, 1)]
                chi = H.chi_outlines_host(parts)[0]
                want = [chi[k] if len(c) >= 20 else small[k] for k, c in enumerate(parts)]
        else:
            want = O.ref_outlines(parts, mode) if O.ref_hull_available() else [w for w, _ in O.convex_outlines(parts, mode)]
        big = [k for k, c in enumerate(parts) if len(c) >= 20]
        assert np.array_equal(host_ids, np.asarray(big if mode == 1 else [], np.uint32))
        at = 0
        for k, c in enumerate(parts):
            got = xy[at:at + int(sizes[k])]
            at += int(sizes[k])
            if mode == 1 and len(c) >= 20:
                assert got.shape[0] == 0
            else:
                assert np.array_equal(got, want[k])
