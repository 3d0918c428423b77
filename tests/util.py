"""Shared helpers for the parity tests (test infrastructure: may import oracle/)."""
from __future__ import annotations

import numpy as np

from oracle.pyoracle import C, J_KOKKOS, J_OPENMP, Oracle  # noqa: F401

import fdtd_method_b200 as fb


def params(Ni, Nj, Nk, dx=C, dy=C, dz=C):
    return fb.Parameters(Ni, Nj, Nk, 0.0, Ni * dx, 0.0, Nj * dy, 0.0, Nk * dz, dx, dy, dz)


def seeded_fields(seed, shape, dtype=np.float64, same_j=True):
    """Same recipe as oracle/make_golden.py: uniform [-1,1], EX..BZ then J."""
    rng = np.random.default_rng(seed)
    f = [rng.uniform(-1.0, 1.0, size=shape) for _ in range(6)]
    j = rng.uniform(-1.0, 1.0, size=shape)
    f += [j, j.copy(), j.copy()] if same_j else [j, rng.uniform(-1, 1, size=shape), rng.uniform(-1, 1, size=shape)]
    return [a.astype(dtype) for a in f]


def make_pair(Ni, Nj, Nk, d=(C, C, C), dt=0.2, dtype=np.float64, pml=None, j_mode=J_KOKKOS, fusion=True,
              pml_thickness=None, **extra):
    """(oracle, gpu solver) on the same grid."""
    o = Oracle(Ni, Nj, Nk, d[0], d[1], d[2], dt, dtype=dtype, j_mode=j_mode, pml_percent=pml, f32_arith=bool(extra.get("f32_arith", False)))
    p = params(Ni, Nj, Nk, *d)
    kw = dict(dtype=dtype, j_openmp_quirk=(j_mode == J_OPENMP), fusion=fusion, **extra)
    if pml is None and pml_thickness is None:
        g = fb.FDTD(p, dt, **kw)
    elif pml_thickness is not None:
        g = fb.FDTD_PML(p, dt, pml_thickness=pml_thickness, **kw)
    else:
        g = fb.FDTD_PML(p, dt, pml, **kw)
    return o, g


def load_both(o, g, fields, comps=range(9)):
    for c in comps:
        o.field(c)[...] = fields[c]
        g.upload(c, fields[c])


def assert_bit_equal(o, g, comps=range(6), what=""):
    for c in comps:
        a, b = o.field(c), g.download(c)
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)
            raise AssertionError(f"{what}: component {c} differs at {len(bad)} cells, first {bad[0]} "
                                 f"oracle={a[tuple(bad[0])]!r} gpu={b[tuple(bad[0])]!r} "
                                 f"max|diff|={np.abs(a.astype(np.float64) - b.astype(np.float64)).max()}")


def rel_linf(a, b):
    """north_star tolerance metric: max|a-b| / max|b| per field."""
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / max(np.abs(b).max(), 1e-300))
