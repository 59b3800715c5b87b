import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_addoption(parser):
    parser.addoption("--qmc-lib", default="", help="run against this build of libqmcnn_b200 (e.g. the debug build with "
                     "device-side bounds checks: make -C qmcnn_b200/csrc debug) instead of the in-tree library")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")
    lib = config.getoption("--qmc-lib")
    if lib:
        from qmcnn_b200 import _lib
        _lib.LIB_PATH = os.path.abspath(lib)


def pytest_report_header(config):
    try:
        from qmcnn_b200 import _lib
        return "libqmcnn_b200: %s  [%s]" % (_lib.LIB_PATH, _lib.load().qmc_version().decode())
    except Exception as e:  # not built yet: the ABI test reports it
        return "libqmcnn_b200: not loaded (%s)" % e


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
