import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        from rolo_b200 import capi
        return capi.load().rg_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests must never silently pass on a CPU box: they are skipped only when no device exists
    # AND the run did not ask for them explicitly with -m gpu.
    if "gpu" in (config.getoption("-m") or ""):
        return
    if not _has_gpu():
        skip = pytest.mark.skip(reason="no CUDA device")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)


@pytest.fixture(scope="session")
def small_pair():
    from rolo_b200 import synth
    return synth.scan_pair("small")          # 32 x 512 = 16384 points per cloud


@pytest.fixture(scope="session")
def tiny_pair():
    from rolo_b200 import synth
    return synth.scan_pair("tiny")           # 16 x 256 = 4096 points per cloud


@pytest.fixture(scope="session")
def c1_pair():
    from rolo_b200 import synth
    return synth.scan_pair("vlp64")          # BASELINE config 1: 64 x 1024


def poses_for_tests():
    from rolo_b200 import synth
    return [np.eye(4), synth.T_GT_DEFAULT, synth.pose((1.0, -2.0, 4.0), (0.15, -0.05, 0.02)), synth.pose((-3.0, 0.5, -2.0), (-0.1, 0.2, -0.05))]
