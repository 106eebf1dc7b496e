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
