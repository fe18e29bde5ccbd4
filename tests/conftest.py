import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_device():
    try:
        import pgr_tk_b200 as pg
        return os.path.exists(pg.library_path()) and pg.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests skip (not fail) on a machine without a CUDA device or without the built library"""
    if any(it.get_closest_marker("gpu") for it in items) and not _have_device():
        skip = pytest.mark.skip(reason="no CUDA device / libpgr_b200.so not built")
        for it in items:
            if it.get_closest_marker("gpu"):
                it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return GOLDEN
