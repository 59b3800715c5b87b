"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/qmcnn_b200.h declares, the ctypes table binds all of them, and - with no GPU - the
product path fails loudly instead of falling back to anything."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "qmcnn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qmc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from qmcnn_b200 import _lib
    names = _declared()
    assert len(names) >= 16
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
        assert n in _lib.SIGNATURES, "ctypes table does not bind %s" % n
    assert sorted(_lib.SIGNATURES) == names
    assert b"sm_100a" in _lib.load().qmc_version()


def test_model_desc_layout_matches_header():
    from qmcnn_b200 import _lib
    # int32 kind, k, n_layers, channels[16], Ly, Lx, reserved[4]
    assert ctypes.sizeof(_lib.ModelDesc) == 4 * (3 + 16 + 2 + 4)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_gpu_means_loud_failure_not_fallback():
    import qmcnn_b200 as q
    from qmcnn_b200 import _lib
    with pytest.raises(q.QmcError):
        q.CRBM(5, 2, 4, 2)
    with pytest.raises(q.QmcError):
        q.DCRBM(3, [8, 8, 8], 2)
    lib = _lib.load()
    d = _lib.ModelDesc()
    d.kind, d.k, d.n_layers, d.Ly, d.Lx = 0, 5, 1, 6, 6
    d.channels[0] = 8
    h = ctypes.c_void_p()
    rc = lib.qmc_create(ctypes.byref(h), 0, ctypes.byref(d))
    assert rc == -4 and not h.value                       # QMC_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.qmc_last_error(None)
    d.k = 4                                               # even filter: rejected before touching the device
    assert lib.qmc_create(ctypes.byref(h), 0, ctypes.byref(d)) == -1


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "qmcnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
