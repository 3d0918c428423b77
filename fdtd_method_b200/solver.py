"""Python mirror of the reference's solver classes on top of the C ABI.

    FDTD(params, dt)                      reference include/FDTD/FDTD.h:8-41, src/FDTD/FDTD.cpp
    FDTD_PML(params, dt, pml_percent)     reference include/FDTD/FDTD_PML.h:10-45, src/FDTD/FDTD_PML.cpp

Same method names and meaning: ``get_field(Component)``, ``update_fields()``, ``zeroed_currents()``.
The reference hands out a mutable ``Field&``; here ``get_field`` returns a host numpy array shaped
[k, j, i] (flat index i + j*Ni + k*Ni*Nj) and ``set_field`` / item assignment through ``FieldView``
writes back, because the data lives in HBM.  All compute happens in libfdtd_b200.so.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _capi
from .structures import Component, Parameters


class FieldView:
    """What ``get_field`` returns: supports ``f[index]`` / ``f[index] = v`` with the reference's flat index
    (sparse scatter/gather on the device) and ``numpy()`` for a dense host copy."""

    def __init__(self, solver: "FDTD", comp: int):
        self._s, self._c = solver, int(comp)

    def __getitem__(self, index):
        idx = np.atleast_1d(np.asarray(index, dtype=np.int64))
        out = self._s.gather(self._c, idx)
        return out[0] if np.isscalar(index) or np.ndim(index) == 0 else out

    def __setitem__(self, index, value):
        idx = np.atleast_1d(np.asarray(index, dtype=np.int64))
        vals = np.broadcast_to(np.asarray(value, dtype=self._s.dtype), idx.shape)
        self._s.scatter(self._c, idx, vals)

    def numpy(self) -> np.ndarray:
        return self._s.download(self._c)

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    def size(self) -> int:
        return self._s.local_cells

    __len__ = size


class FDTD:
    """Periodic Yee solver on one B200 (or one z-slab rank of a multi-GPU ring)."""

    def __init__(self, parameters: Parameters, dt: float, *, dtype=np.float64, j_openmp_quirk: bool = False,
                 fusion: bool = True, temporal: bool = True, overlap: bool = True, pml_split: bool = True, f32_arith: bool = False, uniform_slabs: bool = False,
                 device: int = -1, rank: int = 0, nranks: int = 1,
                 _pml_mode: int = _capi.PML_NONE, _pml_percent: float = 0.0, _pml_thickness=(0, 0, 0)):
        L = _capi.lib()
        cfg = _capi.Config()
        L.fdtd_config_init(ctypes.byref(cfg))
        cfg.grid = parameters
        cfg.dt = float(dt)
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise TypeError("dtype must be float64 or float32")
        cfg.dtype = _capi.F32 if self.dtype == np.float32 else _capi.F64
        cfg.flags = ((_capi.FLAG_J_OPENMP_QUIRK if j_openmp_quirk else 0) | (0 if fusion else _capi.FLAG_NO_FUSION)
                     | (0 if overlap else _capi.FLAG_NO_OVERLAP) | (0 if pml_split else _capi.FLAG_NO_PML_SPLIT)
                     | (0 if temporal else _capi.FLAG_NO_TEMPORAL) | (_capi.FLAG_F32_ARITH if f32_arith else 0)
                     | (_capi.FLAG_UNIFORM_SLABS if uniform_slabs else 0))
        cfg.pml_mode = _pml_mode
        cfg.pml_percent = float(_pml_percent)
        for a in range(3):
            cfg.pml_thickness[a] = int(_pml_thickness[a])
        cfg.device, cfg.rank, cfg.nranks = int(device), int(rank), int(nranks)
        self._h = ctypes.c_void_p()
        _capi.check(L.fdtd_create_ex(ctypes.byref(cfg), ctypes.byref(self._h)))
        self.parameters = parameters
        self.dt = float(dt)
        info = self.info()
        self.k_begin, self.k_end = info.k_begin, info.k_end
        self.local_shape = (info.k_end - info.k_begin, info.Nj, info.Ni)
        self.local_cells = int(np.prod(self.local_shape))

    # ---- the reference's public interface -------------------------------------------------------------
    def get_field(self, this_field) -> FieldView:
        c = int(this_field)
        if c < 0 or c > 8:
            raise LookupError("ERROR: Invalid field component")  # FDTD.cpp:149
        return FieldView(self, c)

    def update_fields(self) -> None:
        _capi.check(_capi.lib().fdtd_update_fields(self._h))

    def zeroed_currents(self) -> None:
        _capi.check(_capi.lib().fdtd_zeroed_currents(self._h))

    # ---- extensions ------------------------------------------------------------------------------------
    def step(self, nsteps: int) -> None:
        _capi.check(_capi.lib().fdtd_step(self._h, int(nsteps)))

    def sync(self) -> None:
        _capi.check(_capi.lib().fdtd_sync(self._h))

    def issue(self) -> None:
        """Issue a recorded update_fields() call without waiting for the device."""
        _capi.check(_capi.lib().fdtd_issue(self._h))

    def flush(self) -> None:
        """Issue deferred work (recorded step, trailing B half step) without waiting for the device."""
        _capi.check(_capi.lib().fdtd_flush(self._h))

    def upload(self, comp, host: np.ndarray) -> None:
        a = np.ascontiguousarray(host, dtype=self.dtype)
        if a.size != self.local_cells:
            raise TypeError(f"expected {self.local_cells} elements, got {a.size}")
        _capi.check(_capi.lib().fdtd_upload(self._h, int(comp), a.ctypes.data, a.size))

    set_field = upload

    def download(self, comp, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.local_shape, dtype=self.dtype)
        assert out.dtype == self.dtype and out.flags.c_contiguous and out.size == self.local_cells
        _capi.check(_capi.lib().fdtd_download(self._h, int(comp), out.ctypes.data, out.size))
        return out

    def scatter(self, comp, idx: np.ndarray, vals: np.ndarray) -> None:
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        vals = np.ascontiguousarray(vals, dtype=self.dtype)
        assert idx.shape == vals.shape
        _capi.check(_capi.lib().fdtd_scatter(self._h, int(comp), idx.ctypes.data, vals.ctypes.data, idx.size))

    def gather(self, comp, idx: np.ndarray) -> np.ndarray:
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        vals = np.zeros(idx.shape, dtype=self.dtype)
        _capi.check(_capi.lib().fdtd_gather(self._h, int(comp), idx.ctypes.data, vals.ctypes.data, idx.size))
        return vals

    def read_slice(self, comp, axis: int, index: int) -> np.ndarray | None:
        """2-D slice at a fixed GLOBAL coordinate along ``axis`` (0 = i, 1 = j, 2 = k), extracted on the device.
        axis 2 -> [Nj, Ni] (None on a rank that does not own the plane); axis 1 -> [nk_local, Ni]; axis 0 -> [nk_local, Nj]."""
        nk, Nj, Ni = self.local_shape
        shape = {2: (Nj, Ni), 1: (nk, Ni), 0: (nk, Nj)}.get(int(axis))
        if shape is None:
            raise TypeError("axis must be 0, 1 or 2")
        out = np.empty(shape, dtype=self.dtype)
        n = ctypes.c_size_t()
        _capi.check(_capi.lib().fdtd_read_slice(self._h, int(comp), int(axis), int(index), out.ctypes.data, out.size, ctypes.byref(n)))
        return out if n.value == out.size else None

    def dump_slices(self, iteration: int, root: str = ".", axis: int = 2, index: int | None = None, components=range(6)) -> list[str]:
        """Write ``<root>/OutFiles_<c+1>/<iteration>.csv`` for each component (Ex=1 ... Bz=6): the ';'-separated
        per-iteration slice files python_script_legend/visualization.py:10-23,34 of the reference reads.  The rank
        that owns the plane writes (axis 2); returns the paths written."""
        import os
        if index is None:
            index = (self.parameters.Ni, self.parameters.Nj, self.parameters.Nk)[axis] // 2
        paths = []
        for c in components:
            a = self.read_slice(c, axis, index)
            if a is None:
                continue
            d = os.path.join(root, f"OutFiles_{int(c) + 1}")
            os.makedirs(d, exist_ok=True)
            path = os.path.join(d, f"{int(iteration)}.csv")
            with open(path, "w") as fh:
                for row in a:
                    fh.write(";".join(repr(float(v)) for v in row) + "\n")
            paths.append(path)
        return paths

    def set_source(self, lo, hi, wx, wy, wz, amp) -> None:
        """Device-resident current source: J = ((amp[t]*wx)*wy)*wz on [lo, hi) before step t."""
        lo_a = (ctypes.c_int * 3)(*[int(v) for v in lo])
        hi_a = (ctypes.c_int * 3)(*[int(v) for v in hi])
        arrs = [np.ascontiguousarray(w, dtype=np.float64) for w in (wx, wy, wz, amp)]
        ptr = [a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) for a in arrs]
        _capi.check(_capi.lib().fdtd_set_source(self._h, lo_a, hi_a, ptr[0], ptr[1], ptr[2], ptr[3], arrs[3].size))

    def clear_source(self) -> None:
        _capi.check(_capi.lib().fdtd_clear_source(self._h))

    def info(self) -> _capi.Info:
        info = _capi.Info()
        _capi.check(_capi.lib().fdtd_get_info(self._h, ctypes.byref(info)))
        return info

    def timer_start(self) -> None:
        _capi.check(_capi.lib().fdtd_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = ctypes.c_double()
        _capi.check(_capi.lib().fdtd_timer_stop(self._h, ctypes.byref(ms)))
        return ms.value

    def timeline_enable(self, max_passes: int) -> None:
        _capi.check(_capi.lib().fdtd_timeline_enable(self._h, int(max_passes)))

    def timeline_read(self, capacity: int = 4096) -> np.ndarray:
        """[n_passes, 4] ms since the first pass start: pass start, halo copies start, halo copies done, pass end."""
        out = np.zeros((capacity, 4), dtype=np.float64)
        n = ctypes.c_int()
        _capi.check(_capi.lib().fdtd_timeline_read(self._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), capacity, ctypes.byref(n)))
        return out[:n.value].copy()

    def comm_init(self, unique_id: bytes) -> None:
        buf = ctypes.create_string_buffer(bytes(unique_id), _capi.NCCL_UNIQUE_ID_BYTES)
        _capi.check(_capi.lib().fdtd_comm_init(self._h, buf, _capi.NCCL_UNIQUE_ID_BYTES))

    def close(self) -> None:
        if getattr(self, "_h", None):
            _capi.lib().fdtd_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class FDTD_PML(FDTD):
    """Split-field PML solver (reference FDTD_PML(Parameters, FP dt, FP pml_percent), FDTD_PML.h:42).
    ``pml_thickness=(pi, pj, pk)`` is the explicit-thickness extension (hard parts)."""

    def __init__(self, parameters: Parameters, dt: float, pml_percent: float | None = None, *, pml_thickness=None, **kw):
        if pml_thickness is not None:
            super().__init__(parameters, dt, _pml_mode=_capi.PML_THICKNESS, _pml_thickness=pml_thickness, **kw)
        else:
            if pml_percent is None:
                raise TypeError("FDTD_PML needs pml_percent or pml_thickness")
            super().__init__(parameters, dt, _pml_mode=_capi.PML_PERCENT, _pml_percent=pml_percent, **kw)


def nccl_unique_id() -> bytes:
    buf = ctypes.create_string_buffer(_capi.NCCL_UNIQUE_ID_BYTES)
    _capi.check(_capi.lib().fdtd_nccl_unique_id(buf, _capi.NCCL_UNIQUE_ID_BYTES))
    return buf.raw


def slab_range(Nk: int, rank: int, nranks: int):
    b, e = ctypes.c_int(), ctypes.c_int()
    _capi.lib().fdtd_slab_range(Nk, rank, nranks, ctypes.byref(b), ctypes.byref(e))
    return b.value, e.value


def pml_profile(N: int, thickness: int, d: float, dt: float):
    sigma, decay, coef2 = (np.zeros(N) for _ in range(3))
    pd = ctypes.POINTER(ctypes.c_double)
    _capi.check(_capi.lib().fdtd_pml_profile(N, thickness, d, dt, sigma.ctypes.data_as(pd),
                                             decay.ctypes.data_as(pd), coef2.ctypes.data_as(pd)))
    return sigma, decay, coef2
