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


def fuzz_call_sequence(seed, make=None, min_nk=4):
    """One seeded random sequence of API calls on (oracle, GPU solver); `make(Ni, Nj, Nk, d, dtype, pml, f32_arith)` builds the
    pair (default: one GPU through make_pair).  See tests/test_parity_gpu.py::test_random_call_sequences_*."""
    import math
    rng = np.random.default_rng(1000 + seed)
    mode = ["f64", "f64", "f32", "f32a"][seed % 4]
    dtype = np.float64 if mode == "f64" else np.float32
    pml = [None, None, 0.15][seed % 3]
    Ni = int(rng.choice([16, 32, 64, 128, 36])); Nj = int(rng.integers(4, 30)); Nk = int(rng.integers(min_nk, max(min_nk + 1, 20)))
    if pml is not None:
        Ni, Nj, Nk = max(Ni, 32), max(Nj, 24), max(Nk, 16)
    if make is None:
        o, g = make_pair(Ni, Nj, Nk, d=(C, 1.25 * C, 0.8 * C), dtype=dtype, pml=pml, f32_arith=(mode == "f32a"))
    else:
        o, g = make(Ni, Nj, Nk, (C, 1.25 * C, 0.8 * C), dtype, pml, mode == "f32a")
    load_both(o, g, seeded_fields(seed, (Nk, Nj, Ni), dtype=dtype, same_j=False), comps=range(6))
    total = Ni * Nj * Nk
    src = None   # (lo, hi, w, amp, t) of an active device source, mirrored on the oracle by hand

    def oracle_step(n):
        nonlocal src
        for _ in range(n):
            if src is not None:
                lo, hi, w, amp, t = src
                if t < len(amp):
                    for k in range(lo[2], hi[2]):
                        for j in range(lo[1], hi[1]):
                            for i in range(lo[0], hi[0]):
                                v = ((amp[t] * w[0][i - lo[0]]) * w[1][j - lo[1]]) * w[2][k - lo[2]]
                                for c in (6, 7, 8):
                                    o.field(c)[k, j, i] = v
                    src = (lo, hi, w, amp, t + 1)
                else:
                    o.zeroed_currents()
                    src = None
            o.update_fields()

    for _ in range(30):
        op = rng.choice(["uf", "uf", "uf", "step", "jbox", "jbox", "jbig", "jone", "ewrite", "bwrite", "gather", "dense", "zero", "source"])
        if op == "uf":
            g.update_fields(); oracle_step(1)
        elif op == "step":
            n = int(rng.integers(1, 6)); g.step(n); oracle_step(n)
        elif op in ("jbox", "jbig", "jone") and src is None:
            if op == "jbig":
                idx = rng.choice(total, size=min(total, 40), replace=False)
            else:
                i0, j0, k0 = int(rng.integers(0, Ni - 2)), int(rng.integers(0, Nj - 2)), int(rng.integers(0, Nk - 2))
                idx = np.array([i + j * Ni + k * Ni * Nj for k in (k0, k0 + 1) for j in (j0, j0 + 1) for i in (i0, i0 + 1)])
            for c in ((6, 7, 8) if op != "jone" else (int(rng.integers(6, 9)),)):
                v = rng.uniform(-1, 1, size=idx.size).astype(dtype)
                g.scatter(c, idx, v); o.field(c).reshape(-1)[idx] = v
        elif op in ("ewrite", "bwrite"):
            c = int(rng.integers(0, 3)) + (3 if op == "bwrite" else 0)
            idx = rng.choice(total, size=5, replace=False)
            v = rng.uniform(-1, 1, size=5).astype(dtype)
            g.scatter(c, idx, v); o.field(c).reshape(-1)[idx] = v
        elif op == "gather":
            c = int(rng.integers(0, 9))
            idx = rng.choice(total, size=7, replace=False)
            assert np.array_equal(g.gather(c, idx), o.field(c).reshape(-1)[idx]), f"seed {seed}: gather comp {c}"
        elif op == "dense":
            assert_bit_equal(o, g, comps=range(9), what=f"seed {seed}: dense read")
        elif op == "zero":
            g.zeroed_currents(); o.zeroed_currents(); src = None
        elif op == "source" and src is None:
            lo = [int(rng.integers(0, N - 3)) for N in (Ni, Nj, Nk)]
            hi = [l + int(rng.integers(1, 4)) for l in lo]
            w = [[0.3 + 0.1 * math.cos(0.7 * i) for i in range(lo[a], hi[a])] for a in range(3)]
            amp = [math.sin(0.4 * (t + 1)) for t in range(int(rng.integers(1, 6)))]
            g.set_source(lo, hi, w[0], w[1], w[2], amp)
            src = (lo, hi, w, amp, 0)
    assert_bit_equal(o, g, comps=range(9), what=f"seed {seed}: final state")
