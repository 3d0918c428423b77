"""ctypes loaders for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

* ``Oracle``    -- oracle/liboracle.so, the from-scratch C restatement
                  (oracle/fdtd_oracle.c).
* ``Reference`` -- oracle/_ref/libfdtd_ref.so, the UNMODIFIED reference sources
                  (src/FDTD/FDTD.cpp, src/FDTD/FDTD_PML.cpp) behind oracle/ref_shim.cpp.
* ``ReferenceKokkos`` -- oracle/_ref/libfdtd_ref_kokkos.so, the UNMODIFIED Kokkos-OpenMP path
                  (src/FDTD_kokkos/*.cpp + the vendored Kokkos) behind oracle/ref_kokkos_shim.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  Nothing under fdtd_method_b200/ does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libfdtd_ref.so")
REF_KOKKOS_SO = os.path.join(HERE, "_ref", "libfdtd_ref_kokkos.so")

# include/Enums.h:5
EX, EY, EZ, BX, BY, BZ, JX, JY, JZ = range(9)
COMPONENTS = ("EX", "EY", "EZ", "BX", "BY", "BZ", "JX", "JY", "JZ")
SPLITS = ("EXY", "EXZ", "EYX", "EYZ", "EZX", "EZY", "BXY", "BXZ", "BYX", "BYZ", "BZX", "BZY")
J_KOKKOS, J_OPENMP = 0, 1

# include/Constants.h:6-11
C = 3e10
PI = 3.14159265358


def build(force: bool = False) -> None:
    """Compile liboracle.so and, when /root/reference is mounted, _ref/libfdtd_ref.so."""
    if force or not os.path.exists(ORACLE_SO) or (
        os.path.getmtime(ORACLE_SO) < max(
            os.path.getmtime(os.path.join(HERE, f)) for f in ("fdtd_oracle.c", "fdtd_oracle_body.inc"))):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if os.path.isdir("/root/reference/src/FDTD") and (force or not os.path.exists(REF_SO)):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
    if os.path.isdir("/root/reference/3rdparty/kokkos/core") and (force or not os.path.exists(REF_KOKKOS_SO)):
        subprocess.call(["make", "-s", "-C", HERE, "ref_kokkos"])   # needs cmake; optional (tests skip without it)


def have_reference() -> bool:
    return os.path.exists(REF_SO)


def have_reference_kokkos() -> bool:
    return os.path.exists(REF_KOKKOS_SO)


class Oracle:
    """The C restatement.  ``pml_percent=None`` -> periodic FDTD; a float -> FDTD_PML."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            build()
            L = ctypes.CDLL(ORACLE_SO)
            L.oracle_create.restype = ctypes.c_void_p
            L.oracle_create.argtypes = [ctypes.c_int] * 3 + [ctypes.c_double] * 4 + [ctypes.c_int] * 2 + [ctypes.c_double]
            L.oracle_field.restype = ctypes.c_void_p
            L.oracle_field.argtypes = [ctypes.c_void_p, ctypes.c_int]
            L.oracle_split.restype = ctypes.c_void_p
            L.oracle_split.argtypes = [ctypes.c_void_p, ctypes.c_int]
            for n in ("oracle_pml_sigma", "oracle_pml_decay", "oracle_pml_coef2"):
                getattr(L, n).restype = ctypes.POINTER(ctypes.c_double)
                getattr(L, n).argtypes = [ctypes.c_void_p, ctypes.c_int]
            L.oracle_pml_size.restype = ctypes.c_int
            L.oracle_pml_size.argtypes = [ctypes.c_void_p, ctypes.c_int]
            L.oracle_coef.restype = ctypes.c_double
            L.oracle_coef.argtypes = [ctypes.c_void_p, ctypes.c_int]
            for n in ("oracle_update_B", "oracle_update_E", "oracle_update_fields",
                      "oracle_zeroed_currents", "oracle_destroy"):
                getattr(L, n).restype = None
                getattr(L, n).argtypes = [ctypes.c_void_p]
            L.oracle_step.restype = None
            L.oracle_step.argtypes = [ctypes.c_void_p, ctypes.c_int]
            cls._lib = L
        return cls._lib

    def __init__(self, Ni, Nj, Nk, dx, dy, dz, dt, dtype=np.float64, j_mode=J_KOKKOS, pml_percent=None, f32_arith=False):
        """f32_arith (float32 only): evaluate the updates in float instead of double -- the restatement of the GPU
        library's opt-in FDTD_FLAG_F32_ARITH mode, not of anything the reference builds."""
        L = self.lib()
        self.shape = (Nk, Nj, Ni)  # numpy view: k slowest, i fastest (index i + j*Ni + k*Ni*Nj)
        self.dtype = np.dtype(dtype)
        self.has_pml = pml_percent is not None
        if f32_arith and self.dtype != np.float32:
            raise TypeError("f32_arith needs dtype=float32")
        self._h = L.oracle_create(Ni, Nj, Nk, dx, dy, dz, dt, (2 if f32_arith else 1) if self.dtype == np.float32 else 0, j_mode,
                                  -1.0 if pml_percent is None else float(pml_percent))
        if not self._h:
            raise ValueError("ERROR: invalid parameters")

    def _view(self, ptr):
        n = self.shape[0] * self.shape[1] * self.shape[2]
        ct = ctypes.c_float if self.dtype == np.float32 else ctypes.c_double
        buf = (ct * n).from_address(ptr)
        return np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def field(self, comp) -> np.ndarray:
        """Mutable numpy view [k, j, i] of the oracle's own storage."""
        return self._view(self.lib().oracle_field(self._h, comp))

    def split(self, which) -> np.ndarray:
        return self._view(self.lib().oracle_split(self._h, which))

    def pml_tables(self, axis):
        n = self.shape[2 - axis]
        L = self.lib()
        return tuple(np.ctypeslib.as_array(f(self._h, axis), shape=(n,)).copy()
                     for f in (L.oracle_pml_sigma, L.oracle_pml_decay, L.oracle_pml_coef2))

    def pml_size(self, axis):
        return self.lib().oracle_pml_size(self._h, axis)

    def coef(self, which):
        return self.lib().oracle_coef(self._h, which)

    def update_B(self):
        self.lib().oracle_update_B(self._h)

    def update_E(self):
        self.lib().oracle_update_E(self._h)

    def update_fields(self):
        self.lib().oracle_update_fields(self._h)

    def step(self, n):
        self.lib().oracle_step(self._h, n)

    def zeroed_currents(self):
        self.lib().oracle_zeroed_currents(self._h)

    def close(self):
        if self._h:
            self.lib().oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Reference:
    """The real reference (FDTD_openmp::FDTD / FDTD_PML) through oracle/ref_shim.cpp.  fp64 only
    (include/FP.h:3; the float build does not compile)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            build()
            if not os.path.exists(REF_SO):
                raise FileNotFoundError(REF_SO)
            L = ctypes.CDLL(REF_SO)
            L.ref_create.restype = ctypes.c_void_p
            L.ref_create.argtypes = [ctypes.c_int] * 3 + [ctypes.c_double] * 11
            L.ref_field.restype = ctypes.c_void_p
            L.ref_field.argtypes = [ctypes.c_void_p, ctypes.c_int]
            for n in ("ref_update_fields", "ref_zeroed_currents", "ref_destroy"):
                getattr(L, n).restype = None
                getattr(L, n).argtypes = [ctypes.c_void_p]
            L.ref_step.restype = None
            L.ref_step.argtypes = [ctypes.c_void_p, ctypes.c_int]
            L.ref_max_threads.restype = ctypes.c_int
            L.ref_set_threads.argtypes = [ctypes.c_int]
            cls._lib = L
        return cls._lib

    def __init__(self, Ni, Nj, Nk, dx, dy, dz, dt, pml_percent=None, box=None):
        L = self.lib()
        self.shape = (Nk, Nj, Ni)
        ax, bx, ay, by, az, bz = box if box is not None else (0.0, Ni * dx, 0.0, Nj * dy, 0.0, Nk * dz)
        self._h = L.ref_create(Ni, Nj, Nk, ax, bx, ay, by, az, bz, dx, dy, dz, dt,
                               -1.0 if pml_percent is None else float(pml_percent))
        if not self._h:
            raise ValueError("ERROR: invalid parameters")

    def field(self, comp) -> np.ndarray:
        ptr = self.lib().ref_field(self._h, comp)
        if not ptr:
            raise RuntimeError("ERROR: Invalid field component")
        n = self.shape[0] * self.shape[1] * self.shape[2]
        buf = (ctypes.c_double * n).from_address(ptr)
        return np.frombuffer(buf, dtype=np.float64).reshape(self.shape)

    def update_fields(self):
        self.lib().ref_update_fields(self._h)

    def step(self, n):
        self.lib().ref_step(self._h, n)

    def zeroed_currents(self):
        self.lib().ref_zeroed_currents(self._h)

    @classmethod
    def max_threads(cls):
        return cls.lib().ref_max_threads()

    @classmethod
    def set_threads(cls, n):
        cls.lib().ref_set_threads(n)

    def close(self):
        if self._h:
            self.lib().ref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ReferenceKokkos:
    """The real reference's Kokkos-OpenMP path (FDTD_kokkos::FDTD / FDTD_PML) through oracle/ref_kokkos_shim.cpp.
    Distinct Jx / Jy / Jz feed Ex / Ey / Ez (kokkos_functors.h:81-89) -- what the oracle's J_KOKKOS mode restates.
    Kokkos is initialised once per process with the OpenMP thread count in effect at the first create()."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            build()
            if not os.path.exists(REF_KOKKOS_SO):
                raise FileNotFoundError(REF_KOKKOS_SO)
            L = ctypes.CDLL(REF_KOKKOS_SO)
            L.refk_create.restype = ctypes.c_void_p
            L.refk_create.argtypes = [ctypes.c_int] * 3 + [ctypes.c_double] * 11
            L.refk_field.restype = ctypes.c_void_p
            L.refk_field.argtypes = [ctypes.c_void_p, ctypes.c_int]
            for n in ("refk_update_fields", "refk_zeroed_currents", "refk_destroy"):
                getattr(L, n).restype = None
                getattr(L, n).argtypes = [ctypes.c_void_p]
            L.refk_step.restype = None
            L.refk_step.argtypes = [ctypes.c_void_p, ctypes.c_int]
            L.refk_threads.restype = ctypes.c_int
            cls._lib = L
        return cls._lib

    def __init__(self, Ni, Nj, Nk, dx, dy, dz, dt, pml_percent=None, box=None):
        L = self.lib()
        self.shape = (Nk, Nj, Ni)
        ax, bx, ay, by, az, bz = box if box is not None else (0.0, Ni * dx, 0.0, Nj * dy, 0.0, Nk * dz)
        self._h = L.refk_create(Ni, Nj, Nk, ax, bx, ay, by, az, bz, dx, dy, dz, dt,
                                -1.0 if pml_percent is None else float(pml_percent))
        if not self._h:
            raise ValueError("ERROR: invalid parameters")

    def field(self, comp) -> np.ndarray:
        ptr = self.lib().refk_field(self._h, comp)
        if not ptr:
            raise RuntimeError("ERROR: Invalid field component")
        n = self.shape[0] * self.shape[1] * self.shape[2]
        buf = (ctypes.c_double * n).from_address(ptr)
        return np.frombuffer(buf, dtype=np.float64).reshape(self.shape)

    def update_fields(self):
        self.lib().refk_update_fields(self._h)

    def step(self, n):
        self.lib().refk_step(self._h, n)

    def zeroed_currents(self):
        self.lib().refk_zeroed_currents(self._h)

    @classmethod
    def threads(cls):
        return cls.lib().refk_threads()

    def close(self):
        if self._h:
            self.lib().refk_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------
# Scenario helpers shared by the tests, the golden generator and bench.py.
# ---------------------------------------------------------------------------
def sample_params(n):
    """Grid of perf-tests/sample/sample.cpp:33-51: dx=dy=dz=C, dt=0.2 (Courant 0.2)."""
    return dict(Ni=n, Nj=n, Nk=n, dx=C, dy=C, dz=C, dt=0.2)


def sample_source(n, iters):
    """Source term of perf-tests/sample/sample.cpp:15-31,57-83, evaluated with
    Python floats in the reference's left-to-right order (same glibc libm).

    Returns (lo, hi, active_steps, value(t, i, j, k))."""
    import math
    T, dt = 8.0, 0.2
    Tx = Ty = Tz = 4.0 * C
    d = C
    a = -(n / 2.0) * d
    lo = tuple(int(math.floor((-Tp / 4.0 - a) / d)) for Tp in (Tx, Ty, Tz))
    hi = tuple(int(math.floor((Tp / 4.0 - a) / d)) for Tp in (Tx, Ty, Tz))
    active = min(int(T / dt), iters)

    def value(t, i, j, k):
        x, y, z, tt = float(i) * d, float(j) * d, float(k) * d, float(t + 1) * dt
        return (math.sin(2.0 * PI * tt / T)
                * math.pow(math.cos(2.0 * PI * x / Tx), 2.0)
                * math.pow(math.cos(2.0 * PI * y / Ty), 2.0)
                * math.pow(math.cos(2.0 * PI * z / Tz), 2.0))

    return lo, hi, active, value


def run_sample(solver, n, iters):
    """Drive ``solver`` (Oracle or Reference) through the sample.cpp scenario."""
    lo, hi, active, value = sample_source(n, iters)
    jx, jy, jz = solver.field(JX), solver.field(JY), solver.field(JZ)
    for t in range(active):
        for k in range(lo[2], hi[2]):
            for j in range(lo[1], hi[1]):
                for i in range(lo[0], hi[0]):
                    v = value(t, i, j, k)
                    jx[k, j, i] = v
                    jy[k, j, i] = v
                    jz[k, j, i] = v
        solver.update_fields()
    solver.zeroed_currents()
    for _ in range(active, iters):
        solver.update_fields()
