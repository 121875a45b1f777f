import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_addoption(parser):
    parser.addoption("--cusim", action="store_true", default=False,
                     help="TEST-SIDE switch: run the -m gpu tests on tests/cusim, the CPU executor for the library's CUDA "
                          "sources (kernel logic against the oracle without a GPU). The product never does this itself.")


CUSIM_SKIP = ("test_gpu_multi.py",)  # two processes with a device each
CUSIM_SKIP_TESTS = ()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config._cusim = bool(config.getoption("--cusim") or os.environ.get("SUMA_B200_TEST_CUSIM") == "1")
    if config._cusim:
        # the swap happens HERE, in the test harness: semantic_suma_b200 has no notion of a CPU executor
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from cusim import build_sim
        path = build_sim.build()
        from semantic_suma_b200 import api, build as product_build
        product_build.LIB = path
        product_build.build = lambda *a, **k: path
        api._lib = None


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
        if getattr(config, "_cusim", False) and (os.path.basename(str(it.fspath)) in CUSIM_SKIP or
                                                  it.name in CUSIM_SKIP_TESTS):
            it.add_marker(pytest.mark.skip(reason="not on the CPU executor"))
            continue
        if "gpu" in it.keywords:
            if gpu is None:
                gpu = _has_gpu()
            if not gpu:
                it.add_marker(pytest.mark.skip(reason="no CUDA device"))
