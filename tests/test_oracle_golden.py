"""Pins the CPU oracle (oracle/fdtd_oracle.c) to the real reference: committed golden vectors produced by
oracle/make_golden.py from the unmodified reference sources, SURVEY.md Appendix B values, and -- when
oracle/_ref/libfdtd_ref.so is present -- the real reference run side by side."""
import json
import math
import os

import numpy as np
import pytest

from oracle.pyoracle import (BX, BY, C, EX, J_KOKKOS, J_OPENMP, Oracle, Reference, have_reference, run_sample,
                             sample_params)
from tests.util import seeded_fields

NAMES = ["EX", "EY", "EZ", "BX", "BY", "BZ"]


@pytest.mark.parametrize("name", ["random_periodic_16x12x10", "random_periodic_33x7x5", "random_pml_20x16x12", "random_pml_24x24x24"])
def test_oracle_matches_golden_random(name, golden_dir):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    m = json.loads(str(z["meta"]))
    o = Oracle(m["Ni"], m["Nj"], m["Nk"], m["dx"], m["dy"], m["dz"], m["dt"], j_mode=J_OPENMP, pml_percent=m["pml_percent"])
    f = seeded_fields(m["seed"], (m["Nk"], m["Nj"], m["Ni"]))
    for c in range(9):
        o.field(c)[...] = f[c]
    done = 0
    for s in m["steps"]:
        o.step(s - done)
        done = s
        for c in range(6):
            assert np.array_equal(o.field(c), z[f"{NAMES[c]}_step{s}"]), f"{name} {NAMES[c]} step {s}"


@pytest.mark.parametrize("pml,name", [(None, "sample_32_100_periodic"), (0.2, "sample_32_100_pml02")])
def test_oracle_matches_golden_sample(pml, name, golden_dir):
    z = np.load(os.path.join(golden_dir, name + "_fields.npz"))
    meta = json.load(open(os.path.join(golden_dir, name + ".json")))
    o = Oracle(**sample_params(32), pml_percent=pml)
    run_sample(o, 32, 100)
    for c in range(6):
        f = o.field(c)
        assert np.array_equal(f, z[NAMES[c]])
        assert float(np.sum(f.astype(np.longdouble) ** 2)) == meta[NAMES[c]]["sumsq"]
        assert float(np.abs(f).max()) == meta[NAMES[c]]["maxabs"]


def test_appendix_b2_values():
    """values extracted from the real reference during the survey."""
    o = Oracle(**sample_params(32))
    assert o.coef(0) == 0.2 and o.coef(3) == 0.1 and o.coef(6) == -2.5132741228640003
    run_sample(o, 32, 100)
    ex, ey, bx, by = o.field(EX), o.field(1), o.field(BX), o.field(BY)
    assert float(np.sum(ex.astype(np.longdouble) ** 2)) == pytest.approx(20.894222322365651, rel=1e-15)
    assert np.abs(ex).max() == 0.1138720861952391
    assert ex[16, 16, 16] == -0.024569547070792515      # point [i,j,k] = [16,16,16]
    assert ex[17, 14, 19] == -0.0053841167475840637     # [19,14,17]
    assert ey[17, 14, 19] == 0.037371595277850306
    assert ex[0, 0, 0] == -4.8024671787795824e-10
    assert ex[1, 16, 31] == -0.0016033115000236571      # [31,16,1]
    assert bx[17, 14, 19] == 0.0074977784938209495
    assert by[1, 16, 31] == 0.0036638471724602514
    assert np.abs(bx).max() == 0.09294634675163338
    # printed slice row 6 of ./sample (B.3)
    assert [f"{v:.5f}" for v in ex[16, 16, 11:21]] == ["0.00373", "-0.02977", "-0.02579", "-0.02357", "-0.02457",
                                                       "-0.02457", "-0.02357", "-0.02579", "-0.02977", "0.00373"]
    p = Oracle(**sample_params(32), pml_percent=0.2)
    assert p.pml_size(0) == 6
    sigma, decay, coef2 = p.pml_tables(0)
    assert sigma[0] == 3.8376418216567427e-10 and sigma[5] == 2.9611433809079797e-13
    assert decay[0] == 0.10000000000000002 and decay[1] == 0.329417687080171
    assert coef2[0] == 0.07817300674258533 and coef2[5] == 0.19982243657087007
    assert np.array_equal(sigma[:6], sigma[::-1][:6])
    run_sample(p, 32, 100)
    ex = p.field(EX)
    assert ex[16, 16, 16] == -0.024407532364765512
    assert ex[0, 0, 0] == -9.807741153588897e-18
    assert p.field(BY)[17, 14, 19] == -0.0053762490766627999
    assert np.abs(ex).max() == 0.11975489771367392


def test_convergence_numbers(golden_dir):
    """unit-tests/test_FDTD_method.cpp ratios from the oracle equal the real reference's err_1/err_2."""
    gold = json.load(open(os.path.join(golden_dir, "convergence.json")))
    # Appendix B.1, 17 significant digits
    assert gold["x_axis_EY"]["err1"] == 0.00060122628834567704
    assert gold["z_axis_BY"]["err2"] == 5.0331679803894414e-05
    PI, T = 3.14159265358, 5e-13
    cases = {"x_axis_EY": (1, 5, 0, 1.0, 1, (16, 8, 4)), "y_axis_BZ": (0, 5, 1, -1.0, 5, (8, 16, 4)),
             "z_axis_EX": (0, 4, 2, 1.0, 0, (4, 8, 16)), "z_axis_BX": (1, 3, 2, -1.0, 3, (4, 8, 16))}

    def run(ef, bf, axis, sign, tf, N):
        Ni, Nj, Nk = N
        d = (1.0 / float(Ni), 2.0 / float(Nj), 3.0 / float(Nk))
        box = (0.0, 1.0, 0.0, 2.0, 0.0, 3.0)
        iters = 16 * (max(N) // 16)
        o = Oracle(Ni, Nj, Nk, d[0], d[1], d[2], T / float(iters))
        a, b = box[2 * axis], box[2 * axis + 1]
        for m in range(N[axis]):
            x = float(m) * d[axis]
            sl = [slice(None)] * 3
            sl[2 - axis] = m
            o.field(ef)[tuple(sl)] = sign * math.sin(2.0 * PI * (x - a) / (b - a))
            o.field(bf)[tuple(sl)] = math.sin(2.0 * PI * (d[axis] / 2.0 + x - a) / (b - a))
        o.step(iters)
        f = o.field(tf)
        is_b = tf > 2
        s, x, err = (1.0 if is_b else sign), (d[axis] / 2.0 if is_b else 0.0), 0.0
        for m in range(N[axis]):
            ix = [0, 0, 0]
            ix[2 - axis] = m
            err = max(err, abs(s * f[tuple(ix)] - math.sin(2.0 * PI * (x - a - C * T) / (b - a))))
            x += d[axis]
        return err

    for name, (ef, bf, ax, sg, tf, N) in cases.items():
        e1, e2 = run(ef, bf, ax, sg, tf, N), run(ef, bf, ax, sg, tf, tuple(2 * v for v in N))
        assert e1 == gold[name]["err1"] and e2 == gold[name]["err2"]
        assert abs(e1 / e2 - 4.0) <= 0.1


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref/libfdtd_ref.so not built (no /root/reference here)")
@pytest.mark.parametrize("shape,pml", [((20, 13, 7), None), ((24, 16, 12), 0.2), ((5, 4, 3), None), ((1, 1, 1), None), ((12, 12, 12), 0.25)])
def test_oracle_vs_live_reference(shape, pml):
    Ni, Nj, Nk = shape
    Reference.set_threads(1)
    o = Oracle(Ni, Nj, Nk, C, 1.1 * C, 0.9 * C, 0.2, j_mode=J_OPENMP, pml_percent=pml)
    r = Reference(Ni, Nj, Nk, C, 1.1 * C, 0.9 * C, 0.2, pml_percent=pml)
    f = seeded_fields(99, (Nk, Nj, Ni), same_j=False)   # distinct J: the OpenMP quirk (G1) must be reproduced
    for c in range(9):
        o.field(c)[...] = f[c]
        r.field(c)[...] = f[c]
    o.step(12)
    r.step(12)
    for c in range(6):
        assert np.array_equal(o.field(c), r.field(c))


def test_kokkos_j_semantics_differ_only_through_jy_jz():
    """J_KOKKOS uses Jy/Jz (kokkos_functors.h:84,87); with Jx=Jy=Jz both modes are bit-identical."""
    f = seeded_fields(4, (6, 7, 8), same_j=True)
    res = []
    for jm in (J_KOKKOS, J_OPENMP):
        o = Oracle(8, 7, 6, C, C, C, 0.2, j_mode=jm)
        for c in range(9):
            o.field(c)[...] = f[c]
        o.step(5)
        res.append([o.field(c).copy() for c in range(6)])
    for c in range(6):
        assert np.array_equal(res[0][c], res[1][c])
    g = seeded_fields(4, (6, 7, 8), same_j=False)
    o = Oracle(8, 7, 6, C, C, C, 0.2, j_mode=J_KOKKOS)
    for c in range(9):
        o.field(c)[...] = g[c]
    o.update_fields()
    assert not np.array_equal(o.field(1), res[0][1])


def test_fp32_oracle_is_float_storage_double_arithmetic():
    o64 = Oracle(8, 8, 8, C, C, C, 0.2)
    o32 = Oracle(8, 8, 8, C, C, C, 0.2, dtype=np.float32)
    f = seeded_fields(1, (8, 8, 8), dtype=np.float32)
    for c in range(9):
        o64.field(c)[...] = f[c]
        o32.field(c)[...] = f[c]
    o64.update_B()
    o32.update_B()
    # one sweep from float-representable inputs: fp32 result == fp64 result rounded once
    for c in (3, 4, 5):
        assert np.array_equal(o32.field(c), o64.field(c).astype(np.float32))


def test_invalid_parameters():
    with pytest.raises(ValueError):
        Oracle(0, 4, 4, C, C, C, 0.2)
    with pytest.raises(ValueError):
        Oracle(4, 4, 4, C, C, C, 0.0)


# ---- the Kokkos path: distinct Jx / Jy / Jz (kokkos_functors.h:81-89) pins the oracle's J_KOKKOS mode -----------------
def test_oracle_j_kokkos_matches_kokkos_golden(golden_dir):
    """Committed outputs of the real FDTD_kokkos::FDTD (oracle/make_golden.py::kokkos_case) with distinct Jx, Jy, Jz."""
    z = np.load(os.path.join(golden_dir, "random_periodic_kokkos_16x12x10.npz"))
    m = json.loads(str(z["meta"]))
    o = Oracle(m["Ni"], m["Nj"], m["Nk"], m["dx"], m["dy"], m["dz"], m["dt"], j_mode=J_KOKKOS)
    f = seeded_fields(m["seed"], (m["Nk"], m["Nj"], m["Ni"]), same_j=False)
    for c in range(9):
        o.field(c)[...] = f[c]
    done = 0
    for s in m["steps"]:
        o.step(s - done)
        done = s
        for c in range(6):
            assert np.array_equal(o.field(c), z[f"{NAMES[c]}_step{s}"]), f"{NAMES[c]} step {s}"
    # and the OpenMP quirk really is a different answer on these inputs (G1)
    q = Oracle(m["Ni"], m["Nj"], m["Nk"], m["dx"], m["dy"], m["dz"], m["dt"], j_mode=J_OPENMP)
    for c in range(9):
        q.field(c)[...] = f[c]
    q.step(1)
    assert not np.array_equal(q.field(1), z["EY_step1"])


def test_oracle_matches_live_kokkos_reference():
    from oracle.pyoracle import ReferenceKokkos, have_reference_kokkos
    if not have_reference_kokkos():
        pytest.skip("oracle/_ref/libfdtd_ref_kokkos.so not built here")
    Ni, Nj, Nk = 24, 10, 14
    d = (C, 1.25 * C, 0.8 * C)
    r = ReferenceKokkos(Ni, Nj, Nk, d[0], d[1], d[2], 0.2)
    o = Oracle(Ni, Nj, Nk, d[0], d[1], d[2], 0.2, j_mode=J_KOKKOS)
    f = seeded_fields(99, (Nk, Nj, Ni), same_j=False)
    for c in range(9):
        r.field(c)[...] = f[c]
        o.field(c)[...] = f[c]
    for _ in range(9):
        r.update_fields()
        o.update_fields()
    for c in range(6):
        assert np.array_equal(r.field(c), o.field(c)), NAMES[c]
    r.zeroed_currents(); o.zeroed_currents()
    r.step(3); o.step(3)
    for c in range(6):
        assert np.array_equal(r.field(c), o.field(c)), NAMES[c]
    r.close()


# ---- randomised side-by-side runs against the live reference builds (skipped where oracle/_ref is absent) ----------------
def test_oracle_matches_live_reference_on_random_small_grids():
    """Property test: random tiny grids (degenerate extents included), spacings, time steps, step counts and PML
    percentages -- the restatement equals FDTD_openmp::FDTD / FDTD_PML bit for bit (Jx feeds all three, G1), and its
    distinct-J mode equals FDTD_kokkos::FDTD on the periodic cases."""
    from hypothesis import given, settings, strategies as st
    from oracle.pyoracle import ReferenceKokkos, have_reference_kokkos
    if not have_reference():
        pytest.skip("oracle/_ref/libfdtd_ref.so not built here")
    Reference.set_threads(1)   # G4: the OpenMP PML has a formal data race

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 9), st.integers(1, 9), st.integers(1, 9), st.sampled_from([None, None, 0.1, 0.2, 0.34]),
           st.integers(1, 5), st.integers(0, 2 ** 31 - 1), st.sampled_from([0.05, 0.2, 0.31]))
    def run(Ni, Nj, Nk, pml, steps, seed, dt):
        rng = np.random.default_rng(seed)
        d = tuple(C * float(v) for v in rng.uniform(0.5, 2.0, size=3))
        f = [rng.uniform(-1, 1, size=(Nk, Nj, Ni)) for _ in range(9)]
        r = Reference(Ni, Nj, Nk, d[0], d[1], d[2], dt, pml_percent=pml)
        o = Oracle(Ni, Nj, Nk, d[0], d[1], d[2], dt, j_mode=J_OPENMP, pml_percent=pml)
        for c in range(9):
            r.field(c)[...] = f[c]
            o.field(c)[...] = f[c]
        r.step(steps); o.step(steps)
        for c in range(6):
            assert np.array_equal(r.field(c), o.field(c)), (Ni, Nj, Nk, pml, steps, NAMES[c])
        r.close()
        if pml is None and have_reference_kokkos():
            k = ReferenceKokkos(Ni, Nj, Nk, d[0], d[1], d[2], dt)
            q = Oracle(Ni, Nj, Nk, d[0], d[1], d[2], dt, j_mode=J_KOKKOS)
            for c in range(9):
                k.field(c)[...] = f[c]
                q.field(c)[...] = f[c]
            k.step(steps); q.step(steps)
            for c in range(6):
                assert np.array_equal(k.field(c), q.field(c)), ("kokkos", Ni, Nj, Nk, steps, NAMES[c])
            k.close()

    run()
