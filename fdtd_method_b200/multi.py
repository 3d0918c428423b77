"""One process, several B200s: the reference's class interface over a z-slab ring of solvers.

The reference's caller has no notion of ranks (include/FDTD/FDTD.h:35-40): it constructs one solver, writes J through
``get_field``, calls ``update_fields()`` and reads fields back.  ``FDTDMulti`` / ``FDTD_PML_Multi`` keep exactly that
interface and spread the grid over ``devices`` as z slabs (coarray/fdtd.F90:149-164): one ``fdtd_solver_t`` per GPU,
linked with ``fdtd_comm_init_local`` (plain peer access: the copy engines push the halo planes, no NCCL, no second
process).  Every call fans out over the slab solvers from this one host thread -- all of them are asynchronous, nothing
blocks on a neighbour on the host; dense reads gather the slabs into one host array, sparse reads / writes take GLOBAL
flat indices like the single-GPU classes.  The C++ twin is ``FDTD_b200::FDTD`` with ``FDTD_B200_GPUS`` /
the device-list constructor (include/FDTD_b200/FDTD.h).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _capi
from .solver import FDTD, FDTD_PML, FieldView


class FDTDMulti:
    _cls = FDTD

    def __init__(self, parameters, dt, *args, devices=None, **kw):
        if devices is None:
            import torch
            devices = list(range(torch.cuda.device_count()))
        self.devices = [int(d) for d in devices]
        n = len(self.devices)
        if n < 1:
            raise ValueError("no devices")
        self.slabs = [self._cls(parameters, dt, *args, device=d, rank=r, nranks=n, **kw) for r, d in enumerate(self.devices)]
        if n > 1:
            arr = (ctypes.c_void_p * n)(*[s._h for s in self.slabs])
            _capi.check(_capi.lib().fdtd_comm_init_local(arr, n))
        self.parameters, self.dt, self.dtype = parameters, float(dt), self.slabs[0].dtype
        self.k_begin, self.k_end = 0, parameters.Nk
        self.local_shape = (parameters.Nk, parameters.Nj, parameters.Ni)
        self.local_cells = int(np.prod(self.local_shape))

    def _prep(self, comp=None, write: bool = False) -> None:
        """Before any call that waits for a slab: every slab issues its recorded work -- a pass of one slab only completes
        once its neighbours have issued theirs.  Accesses that make the library apply the deferred B half step (any access
        to B, and WRITES of E, include/fdtd_b200.h "COLLECTIVE CALLS") get it issued on every slab here, because it needs
        the Ex, Ey ring exchange: left to the per-slab call it would be issued on one slab and waited for at once."""
        c = None if comp is None else int(comp)
        collective = c is not None and (c in (3, 4, 5) or (write and c < 6))
        for s in self.slabs:
            s.flush() if collective else s.issue()

    # ---- the reference's public interface ----------------------------------------------------------------
    def get_field(self, this_field) -> FieldView:
        c = int(this_field)
        if c < 0 or c > 8:
            raise LookupError("ERROR: Invalid field component")
        return FieldView(self, c)

    def update_fields(self) -> None:
        for s in self.slabs:
            s.update_fields()

    def zeroed_currents(self) -> None:
        self._prep()
        for s in self.slabs:
            s.zeroed_currents()

    # ---- extensions, same names as the single-GPU class ------------------------------------------------------
    def step(self, nsteps: int) -> None:
        for s in self.slabs:
            s.step(nsteps)

    def flush(self) -> None:
        for s in self.slabs:
            s.flush()

    def sync(self) -> None:
        for s in self.slabs:       # issue everybody's deferred work first, then wait
            s.flush()
        for s in self.slabs:
            s.sync()

    def upload(self, comp, host: np.ndarray) -> None:
        a = np.ascontiguousarray(host, dtype=self.dtype).reshape(self.local_shape)
        self._prep(comp, write=True)
        for s in self.slabs:
            s.upload(comp, a[s.k_begin:s.k_end])

    set_field = upload

    def download(self, comp, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.local_shape, dtype=self.dtype)
        self._prep(comp)
        for s in self.slabs:
            s.download(comp, out[s.k_begin:s.k_end])
        return out

    def scatter(self, comp, idx, vals) -> None:
        self._prep(comp, write=True)
        for s in self.slabs:       # every slab gets the global list (the J bounding box must agree on all of them)
            s.scatter(comp, idx, vals)

    def gather(self, comp, idx) -> np.ndarray:
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        self._prep(comp)
        plane = self.parameters.Ni * self.parameters.Nj
        out = np.zeros(idx.shape, dtype=self.dtype)
        k = idx // plane
        for s in self.slabs:
            m = (k >= s.k_begin) & (k < s.k_end)
            if m.any():
                out[m] = s.gather(comp, idx[m])
        return out

    def set_source(self, lo, hi, wx, wy, wz, amp) -> None:
        self._prep()
        for s in self.slabs:
            s.set_source(lo, hi, wx, wy, wz, amp)

    def clear_source(self) -> None:
        self._prep()
        for s in self.slabs:
            s.clear_source()

    def info(self):
        return [s.info() for s in self.slabs]

    def close(self) -> None:
        live = [s for s in getattr(self, "slabs", []) if s._h]
        for s in live:
            s.flush()
        for s in live:
            s.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class FDTD_PML_Multi(FDTDMulti):
    _cls = FDTD_PML
