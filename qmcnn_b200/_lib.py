"""ctypes binding of libqmcnn_b200.so (the C ABI in include/qmcnn_b200.h).

PyTorch is used only for device memory and streams; every compute call goes
through the shared library.  There is no CPU or eager fallback: if the library
is missing or no CUDA device is present the calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqmcnn_b200.so")

QMC_MAX_LAYERS = 16
MODEL_CRBM, MODEL_DCRBM = 0, 1
TFIM, HEISENBERG = 0, 1
# tuning / cross-check knobs of a handle (include/qmcnn_b200.h: QMC_FLAG_*, qmc_model_desc.reserved)
FLAG_GENERIC_CONV, FLAG_SWEEP_CLASSIC, FLAG_SWEEP_INPLACE, FLAG_IP_FREE_RUNNING = 1, 2, 4, 8
FLAG_ENERGY_CLASSIC, FLAG_ENERGY_INPLACE, FLAG_BACKWARD_GENERIC, FLAG_IP_ROWMAJOR_SITES = 16, 32, 64, 128
FLAG_FORWARD_BLOCKED, FLAG_BACKWARD_SMEM = 256, 512


class QmcError(RuntimeError):
    pass


class ModelDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("k", C.c_int32), ("n_layers", C.c_int32),
                ("channels", C.c_int32 * QMC_MAX_LAYERS),
                ("Ly", C.c_int32), ("Lx", C.c_int32), ("reserved", C.c_int32 * 4)]


class NdDesc(C.Structure):
    """include/qmcnn_b200.h: qmc_nd_desc (1-D / 3-D lattices)."""
    _fields_ = [("kind", C.c_int32), ("k", C.c_int32), ("n_layers", C.c_int32),
                ("channels", C.c_int32 * QMC_MAX_LAYERS), ("n_dims", C.c_int32),
                ("L", C.c_int32 * 3), ("reserved", C.c_int32 * 3)]


def nd_desc(kind, k, channels, shape):
    d = NdDesc()
    d.kind, d.k, d.n_layers, d.n_dims = kind, k, len(channels), len(shape)
    for i, c in enumerate(channels):
        d.channels[i] = c
    for i, l in enumerate(shape):
        d.L[i] = int(l)
    return d


def check_nd(rc, what):
    if rc != 0:
        msg = load().qmc_nd_last_error()
        raise QmcError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


# name -> (restype, argtypes); kept in one table so tests can check that every
# symbol the header declares is exported and bound.
_vp, _i, _i64, _u64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float, C.c_size_t
SIGNATURES = {
    "qmc_create": (_i, [C.POINTER(_vp), _i, C.POINTER(ModelDesc)]),
    "qmc_destroy": (_i, [_vp]),
    "qmc_last_error": (C.c_char_p, [_vp]),
    "qmc_num_params": (_sz, [_vp]),
    "qmc_receptive_field": (_i, [_vp]),
    "qmc_cache_floats": (_sz, [_vp]),
    "qmc_sweep_workspace_floats": (_sz, [_vp, _i, _i]),
    "qmc_energy_workspace_floats": (_sz, [_vp, _i]),
    "qmc_backward_workspace_floats": (_sz, [_vp, _i]),
    "qmc_set_params": (_i, [_vp, _vp, _vp]),
    "qmc_get_params": (_i, [_vp, _vp, _vp]),
    "qmc_logpsi_forward": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "qmc_metropolis_sweep": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i64, _i64, _vp, _vp, _u64, _i64,
                                  _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp]),
    "qmc_init_spins": (_i, [_i, _vp, _i, _i, _u64, _i64, _i64, _vp]),
    "qmc_set_image_params": (_i, [_vp, _i, _vp, _vp]),
    "qmc_sym_sweep_workspace_floats": (_sz, [_vp, _i, _i, _i]),
    "qmc_metropolis_sweep_sym": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i64, _i64, _vp, _vp, _u64, _i64,
                                      _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp]),
    "qmc_sym_energy_workspace_floats": (_sz, [_vp, _i, _i]),
    "qmc_sym_backward_workspace_floats": (_sz, [_vp, _i, _i]),
    "qmc_logpsi_forward_sym": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "qmc_local_energy_sym": (_i, [_vp, _i, _i, _f, _vp, _i, _vp, _vp, _vp, _vp]),
    "qmc_logpsi_backward_sym": (_i, [_vp, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "qmc_local_energy": (_i, [_vp, _i, _f, _vp, _i, _vp, _vp, _vp, _vp]),
    "qmc_logpsi_backward": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "qmc_nd_last_error": (C.c_char_p, []),
    "qmc_nd_num_params": (_sz, [C.POINTER(NdDesc)]),
    "qmc_nd_scratch_floats": (_sz, [C.POINTER(NdDesc), _i, _i]),
    "qmc_nd_forward": (_i, [C.POINTER(NdDesc), _i, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "qmc_nd_sweep": (_i, [C.POINTER(NdDesc), _i, _vp, _vp, _vp, _vp, _i, _i, _i64, _i64, _vp, _vp, _u64, _i64,
                          _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp]),
    "qmc_nd_local_energy": (_i, [C.POINTER(NdDesc), _i, _i, _f, _vp, _vp, _i, _vp, _vp, _vp]),
    "qmc_nd_backward_scratch_floats": (_sz, [C.POINTER(NdDesc), _i, _i]),
    "qmc_nd_logpsi_backward": (_i, [C.POINTER(NdDesc), _i, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "qmc_diag_peaks": (_i, [_i, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "qmc_diag_peaks2": (_i, [_i, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "qmc_diag_ip_profile": (_i, [C.POINTER(C.c_ulonglong)]),
    "qmc_diag_tanh_check": (_i, [_i, C.POINTER(C.c_ulonglong)]),
    "qmc_diag_sweep_plan": (_i, [C.POINTER(ModelDesc), _i, _i, C.c_int64, _i, C.c_size_t, C.POINTER(C.c_int64)]),
    "qmc_diag_fastdiv_check": (_i, [C.POINTER(C.c_ulonglong)]),
    "qmc_launch_count": (C.c_ulonglong, []),
    "qmc_version": (C.c_char_p, []),
}

_lib = None


def load():
    """Load the shared library (once). Raises QmcError if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise QmcError("%s not found - run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(make -C qmcnn_b200/csrc); there is no CPU fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(handle, rc, what):
    if rc != 0:
        msg = load().qmc_last_error(handle)
        raise QmcError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


class Handle(object):
    """One qmc_handle: a model on one lattice shape on one device."""

    def __init__(self, kind, k, channels, Ly, Lx, device, tuning=None):
        """``tuning``: dict with any of flags (FLAG_* bits), max_warps, ip_group, ip_chunks, ip_stagger -> desc.reserved."""
        lib = load()
        d = ModelDesc()
        d.kind, d.k, d.n_layers, d.Ly, d.Lx = kind, k, len(channels), Ly, Lx
        for i, c in enumerate(channels):
            d.channels[i] = c
        tuning = dict(tuning or {})
        for i, key in enumerate(("flags", "max_warps", "ip_group", "ip_chunks")):
            d.reserved[i] = int(tuning.pop(key, 0))
        if "ip_stagger" in tuning:      # start offset between phase groups, x 1024 cycles (0 = none; absent = default)
            st = int(tuning.pop("ip_stagger"))
            d.reserved[2] = (d.reserved[2] & 0xFF) | ((st if st > 0 else 0xFFFF) << 8)
        if tuning:
            raise QmcError("unknown tuning keys: %s" % sorted(tuning))
        self._h = _vp()
        rc = lib.qmc_create(C.byref(self._h), device, C.byref(d))
        if rc != 0:
            msg = lib.qmc_last_error(None)
            raise QmcError("qmc_create failed (%d): %s" % (rc, msg.decode() if msg else "?"))
        self.device = device
        self.Ly, self.Lx, self.n = Ly, Lx, Ly * Lx
        self.num_params = lib.qmc_num_params(self._h)
        self.cache_floats = lib.qmc_cache_floats(self._h)
        self.r = lib.qmc_receptive_field(self._h)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                load().qmc_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def ptr(self):
        return self._h
