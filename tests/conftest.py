import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) on a machine without a CUDA device: the product has no CPU fallback"""
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as ge

    ge.build()
    return ge.load_package()


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle

    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def golden_frames():
    """The three committed reference frames (0, 77, 153), bit-identical (uint32 words, signed zeros included) to the
    .pcd files: tests/test_oracle_pinning.py::test_frame_cache_is_bit_lossless_against_the_pcd_files."""
    from tools.pack_reference_frames import unpack

    return unpack(ROOT / "tests" / "golden" / "frames_0_77_153.xz")


@pytest.fixture(scope="session")
def fingerprints():
    import json

    return json.loads((ROOT / "tests" / "golden" / "fingerprints.json").read_text())


@pytest.fixture(scope="session")
def synth_small():
    from tests.synth import make_frame

    return make_frame(seed=7, beams=32, azimuth_steps=512)


@pytest.fixture(scope="session")
def ctx(pkg):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    c = pkg.Context(device=0, max_points=300_000, max_frames=8)
    yield c
    c.close()
