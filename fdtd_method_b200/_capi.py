"""ctypes binding of libfdtd_b200.so (the C ABI in include/fdtd_b200.h).

The library is the product; this module only loads it.  If it cannot be loaded the import of the
solver classes fails loudly -- there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import os

from .structures import Parameters

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FDTD_B200_LIB") or os.path.join(_PKG, "libfdtd_b200.so")   # (override: A/B builds, build.py)

F64, F32 = 0, 1
PML_NONE, PML_PERCENT, PML_THICKNESS = 0, 1, 2
FLAG_J_OPENMP_QUIRK, FLAG_NO_FUSION, FLAG_NO_GRAPH, FLAG_NO_OVERLAP = 0x1, 0x2, 0x4, 0x8
FLAG_NO_PML_SPLIT = 0x10
FLAG_NO_TEMPORAL = 0x20
FLAG_UNIFORM_SLABS = 0x40
FLAG_F32_ARITH = 0x80
NCCL_UNIQUE_ID_BYTES = 128

OK, ERR_INVALID_PARAMETERS, ERR_INVALID_COMPONENT, ERR_CUDA, ERR_NCCL, ERR_STATE, ERR_NOMEM, ERR_BAD_ARGUMENT = range(8)


class Config(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_uint32), ("grid", Parameters), ("dt", ctypes.c_double),
                ("dtype", ctypes.c_int32), ("flags", ctypes.c_uint32), ("pml_mode", ctypes.c_int32),
                ("pml_percent", ctypes.c_double), ("pml_thickness", ctypes.c_int32 * 3),
                ("device", ctypes.c_int32), ("rank", ctypes.c_int32), ("nranks", ctypes.c_int32)]


class Info(ctypes.Structure):
    _fields_ = [("Ni", ctypes.c_int32), ("Nj", ctypes.c_int32), ("Nk", ctypes.c_int32),
                ("k_begin", ctypes.c_int32), ("k_end", ctypes.c_int32),
                ("dtype", ctypes.c_int32), ("has_pml", ctypes.c_int32),
                ("pml_thickness", ctypes.c_int32 * 3),
                ("pitch", ctypes.c_int64), ("plane", ctypes.c_int64), ("device_bytes", ctypes.c_int64),
                ("launches", ctypes.c_int64), ("steps_done", ctypes.c_int64),
                ("fused", ctypes.c_int32), ("rank", ctypes.c_int32), ("nranks", ctypes.c_int32),
                ("device", ctypes.c_int32), ("temporal", ctypes.c_int32), ("passes_t2", ctypes.c_int64),
                ("kernel_ns", ctypes.c_int64), ("transport", ctypes.c_int32), ("halo_in_kernel", ctypes.c_int32),
                ("f32_arith", ctypes.c_int32), ("reserved0", ctypes.c_int32)]


# every entry point include/fdtd_b200.h declares: name -> (restype, argtypes)
_vp, _i, _d, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_size_t
_pd = ctypes.POINTER(ctypes.c_double)
_pi = ctypes.POINTER(ctypes.c_int)
SIGNATURES = {
    "fdtd_create": (_i, [ctypes.POINTER(Parameters), _d, ctypes.POINTER(_vp)]),
    "fdtd_create_pml": (_i, [ctypes.POINTER(Parameters), _d, _d, ctypes.POINTER(_vp)]),
    "fdtd_create_ex": (_i, [ctypes.POINTER(Config), ctypes.POINTER(_vp)]),
    "fdtd_config_init": (None, [ctypes.POINTER(Config)]),
    "fdtd_destroy": (_i, [_vp]),
    "fdtd_update_fields": (_i, [_vp]),
    "fdtd_step": (_i, [_vp, _i]),
    "fdtd_zeroed_currents": (_i, [_vp]),
    "fdtd_upload": (_i, [_vp, _i, _vp, _sz]),
    "fdtd_download": (_i, [_vp, _i, _vp, _sz]),
    "fdtd_scatter": (_i, [_vp, _i, _vp, _vp, _sz]),
    "fdtd_gather": (_i, [_vp, _i, _vp, _vp, _sz]),
    "fdtd_read_slice": (_i, [_vp, _i, _i, _i, _vp, _sz, ctypes.POINTER(_sz)]),
    "fdtd_set_source": (_i, [_vp, _pi, _pi, _pd, _pd, _pd, _pd, _i]),
    "fdtd_clear_source": (_i, [_vp]),
    "fdtd_issue": (_i, [_vp]),
    "fdtd_flush": (_i, [_vp]),
    "fdtd_sync": (_i, [_vp]),
    "fdtd_device_ptr": (_i, [_vp, _i, ctypes.POINTER(_vp)]),
    "fdtd_get_info": (_i, [_vp, ctypes.POINTER(Info)]),
    "fdtd_timer_start": (_i, [_vp]),
    "fdtd_timer_stop": (_i, [_vp, _pd]),
    "fdtd_get_stream": (_i, [_vp, ctypes.POINTER(_vp)]),
    "fdtd_nccl_unique_id": (_i, [_vp, _sz]),
    "fdtd_comm_init": (_i, [_vp, _vp, _sz]),
    "fdtd_comm_init_local": (_i, [ctypes.POINTER(_vp), _i]),
    "fdtd_timeline_enable": (_i, [_vp, _i]),
    "fdtd_timeline_read": (_i, [_vp, _pd, _i, _pi]),
    "fdtd_slab_range": (None, [_i, _i, _i, _pi, _pi]),
    "fdtd_slab_range_cfg": (None, [ctypes.POINTER(Config), _i, _pi, _pi]),
    "fdtd_pml_profile": (_i, [_i, _i, _d, _d, _pd, _pd, _pd]),
    "fdtd_pml_thickness": (_i, [_i, _d]),
    "fdtd_debug_t2_chunk_plan": (_i, [_i] * 9 + [_pi, _pi, _i]),
    "fdtd_last_error": (ctypes.c_char_p, []),
    "fdtd_version": (_i, []),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load libfdtd_b200.so (building it with nvcc first if the in-tree binary is missing or stale)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) or os.environ.get("FDTD_B200_REBUILD"):
            from . import build as _build
            _build.build(force=bool(os.environ.get("FDTD_B200_REBUILD")))
        try:
            L = ctypes.CDLL(LIB_PATH)
        except OSError as e:  # no fallback: the CUDA library IS the implementation
            raise ImportError(f"cannot load {LIB_PATH}: {e}. fdtd_method_b200 has no CPU fallback; "
                              f"build it with `python -m fdtd_method_b200.build`.") from e
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return lib().fdtd_last_error().decode()


def check(status: int) -> None:
    """Map C status codes onto the reference's exception types."""
    if status == OK:
        return
    msg = last_error()
    if status == ERR_INVALID_PARAMETERS:
        raise ValueError(msg)        # std::invalid_argument, src/FDTD/FDTD.cpp:5-7
    if status == ERR_INVALID_COMPONENT:
        raise LookupError(msg)       # std::logic_error, src/FDTD/FDTD.cpp:149
    if status == ERR_NOMEM:
        raise MemoryError(msg)
    if status == ERR_BAD_ARGUMENT:
        raise TypeError(msg)
    raise RuntimeError(f"fdtd_b200 status {status}: {msg}")
