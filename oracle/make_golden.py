"""Golden-vector generator.  TEST INFRASTRUCTURE ONLY.

Runs the UNMODIFIED reference (oracle/_ref/libfdtd_ref.so, built by oracle/Makefile from
/root/reference/src/FDTD/*.cpp) on seeded inputs and writes small fixtures under tests/golden/.
Must be run in the container where /root/reference is mounted:

    python oracle/make_golden.py

The fixtures travel with the repository; the GPU box never needs /root/reference.
Each fixture stores the inputs' recipe (seed, grid, steps) and the reference's outputs.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.pyoracle import (BX, BY, BZ, C, COMPONENTS, EX, EY, EZ, JX, JY, JZ, PI, Reference,  # noqa: E402
                             ReferenceKokkos, have_reference_kokkos, run_sample, sample_params)

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def seeded_fields(seed, shape, same_j=True):
    """Uniform [-1,1] fields in component order EX..BZ, then J (Jx=Jy=Jz when same_j)."""
    rng = np.random.default_rng(seed)
    f = [rng.uniform(-1.0, 1.0, size=shape) for _ in range(6)]
    j = rng.uniform(-1.0, 1.0, size=shape)
    f += [j, j.copy(), j.copy()] if same_j else [j, rng.uniform(-1, 1, size=shape), rng.uniform(-1, 1, size=shape)]
    return f


def random_case(name, Ni, Nj, Nk, d, dt, steps, seed, pml):
    Reference.set_threads(1)  # PML has a formal data race; pin golden at 1 thread
    r = Reference(Ni, Nj, Nk, d[0], d[1], d[2], dt, pml_percent=pml)
    init = seeded_fields(seed, (Nk, Nj, Ni))
    for c in range(9):
        r.field(c)[...] = init[c]
    snaps = {}
    done = 0
    for s in steps:
        r.step(s - done)
        done = s
        for c in range(6):
            snaps[f"{COMPONENTS[c]}_step{s}"] = r.field(c).copy()
    meta = dict(Ni=Ni, Nj=Nj, Nk=Nk, dx=d[0], dy=d[1], dz=d[2], dt=dt, steps=list(steps), seed=seed,
                pml_percent=pml, rng="numpy default_rng(seed).uniform(-1,1) EX..BZ then J (Jx=Jy=Jz)",
                source="oracle/_ref/libfdtd_ref.so (FDTD_openmp, g++ -O3 -fopenmp -DNDEBUG, 1 thread)")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(meta), **snaps)
    print("wrote", name, {k: float(np.abs(v).max()) for k, v in list(snaps.items())[:2]})


def kokkos_case(name, Ni, Nj, Nk, d, dt, steps, seed, pml):
    """Distinct Jx / Jy / Jz through the real Kokkos path (kokkos_functors.h:81-89): pins the oracle's J_KOKKOS mode.
    (Kokkos is initialised with whatever OMP_NUM_THREADS is in effect; __main__ re-executes itself with
    OMP_NUM_THREADS=1 for these cases.)"""
    r = ReferenceKokkos(Ni, Nj, Nk, d[0], d[1], d[2], dt, pml_percent=pml)
    init = seeded_fields(seed, (Nk, Nj, Ni), same_j=False)
    for c in range(9):
        r.field(c)[...] = init[c]
    snaps = {}
    done = 0
    for s in steps:
        r.step(s - done)
        done = s
        for c in range(6):
            snaps[f"{COMPONENTS[c]}_step{s}"] = r.field(c).copy()
    meta = dict(Ni=Ni, Nj=Nj, Nk=Nk, dx=d[0], dy=d[1], dz=d[2], dt=dt, steps=list(steps), seed=seed,
                pml_percent=pml, rng="numpy default_rng(seed).uniform(-1,1) EX..BZ then Jx, Jy, Jz (distinct)",
                source=f"oracle/_ref/libfdtd_ref_kokkos.so (FDTD_kokkos, Kokkos OpenMP backend, g++ -O3 -fopenmp -DNDEBUG, {ReferenceKokkos.threads()} thread(s))")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(meta), **snaps)
    print("wrote", name, {k: float(np.abs(v).max()) for k, v in list(snaps.items())[:2]})


def sample_case(name, n, iters, pml):
    Reference.set_threads(1)
    p = sample_params(n)
    r = Reference(p["Ni"], p["Nj"], p["Nk"], p["dx"], p["dy"], p["dz"], p["dt"], pml_percent=pml)
    run_sample(r, n, iters)
    out = {}
    pts = [(16, 16, 16), (19, 14, 17), (0, 0, 0), (31, 16, 1)]  # (i, j, k), SURVEY.md B.2
    for c in range(6):
        f = r.field(c)
        out[COMPONENTS[c]] = dict(
            sumsq=float(np.sum(f.astype(np.longdouble) ** 2)),
            maxabs=float(np.abs(f).max()),
            points={f"{i},{j},{k}": float(f[k, j, i]) for (i, j, k) in pts if max(i, j, k) < n},
        )
    k = n // 2
    out["slice_EX_xy"] = r.field(EX)[k, n // 2 - 5:n // 2 + 5, n // 2 - 5:n // 2 + 5].tolist()
    out["slice_EX_yz"] = r.field(EX)[n // 2 - 5:n // 2 + 5, n // 2 - 5:n // 2 + 5, n // 2].T.tolist()
    out["meta"] = dict(n=n, iters=iters, pml_percent=pml, scenario="perf-tests/sample/sample.cpp spherical_wave")
    with open(os.path.join(OUT, name + ".json"), "w") as fh:
        json.dump(out, fh, indent=1)
    np.savez_compressed(os.path.join(OUT, name + "_fields.npz"), **{COMPONENTS[c]: r.field(c) for c in range(6)})
    print("wrote", name, out["EX"]["sumsq"])


def convergence_case(name):
    """err_1, err_2 of unit-tests/test_FDTD_method.cpp:18-69 for the six distinct tests, computed by
    driving the real reference with the fixture arithmetic of src/FDTD/test_FDTD.cpp:5-51,89-130."""
    import math
    Reference.set_threads(1)
    T = 5e-13

    def run(efield, bfield, axis, sign, test_field, N):
        Ni, Nj, Nk = N
        box = (0.0, 1.0, 0.0, 2.0, 0.0, 3.0)
        d = (1.0 / float(Ni), 2.0 / float(Nj), 3.0 / float(Nk))
        iters = 16 * (max(N) // 16)
        dt = T / float(iters)
        r = Reference(Ni, Nj, Nk, d[0], d[1], d[2], dt, box=box)
        a, b = box[2 * axis], box[2 * axis + 1]
        n_ax = N[axis]
        e, bb = r.field(efield), r.field(bfield)
        for m in range(n_ax):
            x = float(m) * d[axis]
            ve = sign * math.sin(2.0 * PI * (x - a) / (b - a))
            vb = math.sin(2.0 * PI * (d[axis] / 2.0 + x - a) / (b - a))
            idx = [slice(None)] * 3
            idx[2 - axis] = m
            e[tuple(idx)] = ve
            bb[tuple(idx)] = vb
        r.step(iters)
        f = r.field(test_field)
        is_b = test_field > EZ
        s = 1.0 if is_b else sign
        x = d[axis] / 2.0 if is_b else 0.0
        err = 0.0
        for m in range(n_ax):
            idx = [0, 0, 0]
            idx[2 - axis] = m
            v = f[tuple(idx)]
            err = max(err, abs(s * v - math.sin(2.0 * PI * (x - a - C * T) / (b - a))))
            x += d[axis]
        return err

    cases = {
        "x_axis_EY": (EY, BZ, 0, 1.0, EY, (16, 8, 4)), "x_axis_BZ": (EY, BZ, 0, 1.0, BZ, (16, 8, 4)),
        "x_axis_EZ": (EZ, BY, 0, -1.0, EZ, (16, 8, 4)), "x_axis_BY": (EZ, BY, 0, -1.0, BY, (16, 8, 4)),
        "y_axis_EX": (EX, BZ, 1, -1.0, EX, (8, 16, 4)), "y_axis_BZ": (EX, BZ, 1, -1.0, BZ, (8, 16, 4)),
        "y_axis_EZ": (EZ, BX, 1, 1.0, EZ, (8, 16, 4)), "y_axis_BX": (EZ, BX, 1, 1.0, BX, (8, 16, 4)),
        "z_axis_EX": (EX, BY, 2, 1.0, EX, (4, 8, 16)), "z_axis_BY": (EX, BY, 2, 1.0, BY, (4, 8, 16)),
        "z_axis_EY": (EY, BX, 2, -1.0, EY, (4, 8, 16)), "z_axis_BX": (EY, BX, 2, -1.0, BX, (4, 8, 16)),
    }
    out = {}
    for k, (e, b, ax, sg, tf, N) in cases.items():
        e1 = run(e, b, ax, sg, tf, N)
        e2 = run(e, b, ax, sg, tf, tuple(2 * v for v in N))
        out[k] = dict(err1=e1, err2=e2, ratio=e1 / e2)
        print(k, e1, e2, e1 / e2)
    with open(os.path.join(OUT, name + ".json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "--kokkos-only" in sys.argv or have_reference_kokkos():
        if os.environ.get("OMP_NUM_THREADS") != "1":   # pin the Kokkos goldens at one thread (G4), in a fresh process
            import subprocess
            subprocess.check_call([sys.executable, os.path.abspath(__file__), "--kokkos-only"], env=dict(os.environ, OMP_NUM_THREADS="1"))
        else:
            kokkos_case("random_periodic_kokkos_16x12x10", 16, 12, 10, (C, 1.25 * C, 0.8 * C), 0.2, (1, 10, 30), 45, None)
            # (no Kokkos PML golden: FDTD_PML_kokkos.cpp:26-45 allocates the split fields WithoutInitializing and never
            # zero-fills them, SURVEY.md G5 -- the run produced NaN here; the OpenMP PML is the PML oracle)
            sys.exit(0)
    if "--kokkos-only" in sys.argv:
        sys.exit(0)
    random_case("random_periodic_16x12x10", 16, 12, 10, (C, 1.25 * C, 0.8 * C), 0.2, (1, 10, 50), 42, None)
    random_case("random_periodic_33x7x5", 33, 7, 5, (C, C, C), 0.2, (3, 20), 7, None)
    random_case("random_pml_20x16x12", 20, 16, 12, (C, 1.25 * C, 0.8 * C), 0.2, (1, 10, 40), 43, 0.2)
    random_case("random_pml_24x24x24", 24, 24, 24, (C, C, C), 0.2, (25,), 44, 0.13)
    sample_case("sample_32_100_periodic", 32, 100, None)
    sample_case("sample_32_100_pml02", 32, 100, 0.2)
    convergence_case("convergence")
