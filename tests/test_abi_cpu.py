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


def test_every_export_survives_null_arguments():
    """The C ABI never throws, aborts or dereferences a null pointer: every entry point called with NULL pointers and zero
    sizes returns (an error code, or 0 for the size queries).  In a subprocess: a crash must fail this test, not pytest."""
    import subprocess
    import sys
    code = r"""
import ctypes as C, sys
sys.path.insert(0, %r)
from qmcnn_b200 import _lib
lib = _lib.load()
ints = (C.c_int, C.c_int32, C.c_int64, C.c_uint64, C.c_size_t, C.c_longlong, C.c_ulonglong)
for n in sorted(_lib.SIGNATURES):
    res, args = _lib.SIGNATURES[n]
    vals = [a(0) if a in ints else a(0.0) if a in (C.c_float, C.c_double) else None for a in args]
    print("call", n, flush=True)
    r = getattr(lib, n)(*vals)
    if res is C.c_int and n not in ("qmc_destroy", "qmc_receptive_field"):      # (destroy(NULL) is a no-op; r is a query)
        assert r < 0, (n, r)          # a null handle / descriptor / output is an error, never success
    elif res in (C.c_size_t, C.c_int):
        assert r == 0, (n, r)         # size / shape queries answer 0 without a handle
print("survived", len(_lib.SIGNATURES))
""" % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    last = [l for l in r.stdout.splitlines() if l.startswith("call")][-1:]
    assert r.returncode == 0 and "survived" in r.stdout, "crashed or failed at %s: %s" % (last, r.stderr[-1500:])


def test_fastdiv_magics_divide_exactly():
    """Index arithmetic of every kernel: x / d as one multiply-high (FastDiv), magics of d < 512 from the constant-memory
    table.  qmc_diag_fastdiv_check compares it with integer division on the host (same constexpr table)."""
    from qmcnn_b200 import _lib
    bad = ctypes.c_ulonglong(1)
    assert _lib.load().qmc_diag_fastdiv_check(ctypes.byref(bad)) == 0
    assert bad.value == 0
