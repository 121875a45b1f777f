import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _has_gpu():
    try:
        from semantic_suma_b200 import api
        return api.lib().sb_device_count() > 0
    except Exception:  # noqa: BLE001
        return False


@pytest.fixture(scope="session")
def has_gpu():
    return _has_gpu()


def pytest_collection_modifyitems(config, items):
    # GPU tests must run on the CUDA path; on a box without a GPU they are skipped (the driver deselects them with -m)
    gpu = None
    for it in items:
        if "gpu" in it.keywords:
            if gpu is None:
                gpu = _has_gpu()
            if not gpu:
                it.add_marker(pytest.mark.skip(reason="no CUDA device"))
