"""GPU parity: libfdtd_b200.so (through the C ABI) against the CPU oracle and the committed golden
vectors of the real reference.  Bar (north_star): bit-exact in fp64 and in fp32-storage mode (the
kernels never contract to FMA); the stated tolerances 1e-12 / 1e-5 relative L-inf are asserted as well.
"""
import json
import os

import numpy as np
import pytest

from oracle.pyoracle import C, J_KOKKOS, J_OPENMP, Oracle, run_sample, sample_params, sample_source
from tests.util import assert_bit_equal, load_both, make_pair, params, rel_linf, seeded_fields

import fdtd_method_b200 as fb

pytestmark = pytest.mark.gpu

TOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}


@pytest.mark.parametrize("fusion", [True, False])
@pytest.mark.parametrize("shape,steps", [((16, 12, 10), 7), ((64, 64, 64), 5), ((32, 8, 4), 9), ((2, 1, 3), 4),
                                         ((62, 9, 5), 3), ((60, 6, 5), 3), ((128, 20, 40), 4), ((4, 4, 70), 3)])
def test_periodic_random_bit_exact(shape, steps, fusion):
    Ni, Nj, Nk = shape
    o, g = make_pair(Ni, Nj, Nk, d=(C, 1.25 * C, 0.8 * C), fusion=fusion)
    assert bool(g.info().fused) == fusion
    load_both(o, g, seeded_fields(11, (Nk, Nj, Ni), same_j=False))
    o.step(steps)
    g.step(steps)
    assert_bit_equal(o, g, what=f"periodic {shape} fusion={fusion}")
    for c in range(6):
        assert rel_linf(g.download(c), o.field(c)) <= 1e-12


@pytest.mark.parametrize("shape", [(33, 7, 5), (1, 1, 1), (3, 5, 2), (17, 3, 9)])
def test_periodic_odd_sizes_use_sweeps(shape):
    """Ni not a multiple of the vector width: the two-sweep kernels take over (still CUDA)."""
    Ni, Nj, Nk = shape
    o, g = make_pair(Ni, Nj, Nk)
    load_both(o, g, seeded_fields(5, (Nk, Nj, Ni), same_j=False))
    o.step(6)
    g.step(6)
    assert_bit_equal(o, g, what=f"odd {shape}")


def test_update_fields_one_by_one_equals_step_n():
    """fdtd_step(n) must be bit-identical to n x update_fields() with reads in between (deferred half step)."""
    Ni, Nj, Nk = 32, 16, 12
    o, g = make_pair(Ni, Nj, Nk)
    f = seeded_fields(3, (Nk, Nj, Ni))
    load_both(o, g, f)
    for t in range(5):
        o.update_fields()
        g.update_fields()
        if t % 2 == 0:
            assert_bit_equal(o, g, what=f"step {t}")   # forces a flush of the pending half step
    assert_bit_equal(o, g, what="final")


def test_update_fields_loop_pairs_lazily(monkeypatch):
    """Reference-style caller loops (`for (...) solver.update_fields();`) reach the two-step pass: an odd call is
    recorded, the next one issues both steps; any state access in between runs the recorded step first."""
    Ni, Nj, Nk = 32, 16, 12
    o, g = make_pair(Ni, Nj, Nk)
    f = seeded_fields(4, (Nk, Nj, Ni), same_j=False)
    load_both(o, g, f)
    for _ in range(7):
        o.update_fields()
        g.update_fields()
    info = g.info()
    assert info.steps_done == 7 and info.passes_t2 == 3     # three pairs issued, the 7th call recorded
    assert_bit_equal(o, g, what="7 calls")                   # the download runs the recorded step
    assert g.info().passes_t2 == 3
    # a J write between two calls: the recorded step must still see the old J
    o.update_fields(); g.update_fields()
    idx = np.array([5, 77, 300, Ni * Nj * Nk - 1])
    vals = np.array([0.5, -1.5, 2.0, 3.0])
    for c in (6, 7, 8):
        g.scatter(c, idx, vals * (c - 5))
        o.field(c).reshape(-1)[idx] = vals * (c - 5)
    o.update_fields(); g.update_fields()
    o.update_fields(); g.update_fields()
    assert_bit_equal(o, g, what="J write between calls")
    # switched off: every call is issued at once, same bits
    monkeypatch.setenv("FDTD_B200_NO_LAZY", "1")
    o2, g2 = make_pair(Ni, Nj, Nk)
    load_both(o2, g2, f)
    for _ in range(4):
        o2.update_fields(); g2.update_fields()
    assert g2.info().passes_t2 == 0
    assert_bit_equal(o2, g2, what="no lazy")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("fusion", [True, False])
def test_dtype_modes(dtype, fusion):
    Ni, Nj, Nk = 64, 24, 16
    o, g = make_pair(Ni, Nj, Nk, dtype=dtype, fusion=fusion)
    load_both(o, g, seeded_fields(21, (Nk, Nj, Ni), dtype=dtype, same_j=False))
    o.step(10)
    g.step(10)
    assert_bit_equal(o, g, what=f"{dtype} fusion={fusion}")
    for c in range(6):
        assert rel_linf(g.download(c), o.field(c)) <= TOL[np.dtype(dtype)]


def test_fp32_storage_tracks_fp64_within_tolerance():
    """north_star: <= 1e-5 relative L-inf in fp32 after N steps (against the fp64 reference semantics)."""
    n, steps = 32, 100
    o = Oracle(**sample_params(n))
    run_sample(o, n, steps)
    g = fb.FDTD(params(n, n, n), 0.2, dtype=np.float32)
    lo, hi, active, value = sample_source(n, steps)
    idx = np.array([i + j * n + k * n * n for k in range(lo[2], hi[2]) for j in range(lo[1], hi[1]) for i in range(lo[0], hi[0])])
    for t in range(active):
        vals = np.array([value(t, i % n, (i // n) % n, i // (n * n)) for i in idx], dtype=np.float32)
        for c in (6, 7, 8):
            g.scatter(c, idx, vals)
        g.update_fields()
    g.zeroed_currents()
    g.step(steps - active)
    for c in range(6):
        assert rel_linf(g.download(c), o.field(c)) <= 1e-5


@pytest.mark.parametrize("j_mode", [J_KOKKOS, J_OPENMP])
@pytest.mark.parametrize("fusion", [True, False])
def test_current_semantics(j_mode, fusion):
    """Distinct Jx/Jy/Jz: Kokkos semantics by default, FDTD_openmp's Jx-for-all quirk behind the flag (G1)."""
    Ni, Nj, Nk = 32, 12, 9
    o, g = make_pair(Ni, Nj, Nk, j_mode=j_mode, fusion=fusion)
    load_both(o, g, seeded_fields(8, (Nk, Nj, Ni), same_j=False))
    o.step(4)
    g.step(4)
    assert_bit_equal(o, g, what=f"j_mode={j_mode}")


@pytest.mark.parametrize("shape,pml,steps", [((20, 16, 12), 0.2, 10), ((24, 24, 24), 0.13, 6), ((32, 32, 32), 0.2, 12),
                                            ((33, 10, 10), 0.1, 5), ((16, 16, 16), 0.0, 4), ((8, 8, 8), 0.5, 3)])
def test_pml_random_bit_exact(shape, pml, steps):
    Ni, Nj, Nk = shape
    o, g = make_pair(Ni, Nj, Nk, d=(C, 1.25 * C, 0.8 * C), pml=pml)
    load_both(o, g, seeded_fields(43, (Nk, Nj, Ni)))
    for t in range(steps):
        o.update_fields()
        g.update_fields()
        if t == 1:
            assert_bit_equal(o, g, what=f"pml {shape} step {t}")
    assert_bit_equal(o, g, what=f"pml {shape}")


@pytest.mark.parametrize("pml_split", [True, False])
@pytest.mark.parametrize("shape,pml,dtype", [((64, 48, 48), 0.1, np.float64), ((50, 44, 44), 0.1, np.float64),
                                             ((52, 44, 44), 0.1, np.float32), ((64, 40, 72), 0.07, np.float64)])
def test_pml_interior_shell_split(shape, pml, dtype, pml_split):
    """Grids large enough for the two-launch PML path (lean kernel on the aligned interior box, PML kernel on
    the shell and the unaligned fringe) -- must equal the oracle and the single-launch path bit for bit."""
    Ni, Nj, Nk = shape
    o, g = make_pair(Ni, Nj, Nk, d=(C, 1.25 * C, 0.8 * C), pml=pml, dtype=dtype, pml_split=pml_split)
    load_both(o, g, seeded_fields(77, (Nk, Nj, Ni), dtype=dtype, same_j=False))
    for t in range(6):
        o.update_fields()
        g.update_fields()
        if t == 2:
            assert_bit_equal(o, g, what=f"pml split={pml_split} {shape} mid-run (flush path)")
    assert_bit_equal(o, g, what=f"pml split={pml_split} {shape}")


@pytest.mark.parametrize("shape,pml,dtype", [((64, 48, 48), 0.1, np.float64), ((50, 44, 44), 0.1, np.float64),
                                             ((52, 44, 44), 0.1, np.float32), ((64, 9, 48), 0.1, np.float64),
                                             ((64, 40, 72), 0.07, np.float64), ((128, 64, 40), 0.2, np.float64),
                                             ((32, 64, 9), 0.1, np.float64)])
def test_pml_two_step_pass(shape, pml, dtype, monkeypatch):
    """fdtd_step(n >= 2) on a PML solver: T2 pass on the main box's core + rim sweeps for the shell and the band
    around it (two steps per pass) must equal n x update_fields() of the oracle bit for bit -- distinct Jx/Jy/Jz
    everywhere (so the rim reads J), thickness 0 on an axis (periodic there), fp32, odd step counts, reads in between."""
    Ni, Nj, Nk = shape
    monkeypatch.setenv("FDTD_B200_PML_T2_F32", "1")   # the fp32 pair path is off by default (slower than the sweeps)
    o, g = make_pair(Ni, Nj, Nk, d=(C, 1.25 * C, 0.8 * C), pml=pml, dtype=dtype)
    assert g.info().temporal == 1 and g.info().fused == 0
    load_both(o, g, seeded_fields(91, (Nk, Nj, Ni), dtype=dtype, same_j=False))
    o.step(7)
    g.step(7)
    assert g.info().passes_t2 == 3
    assert_bit_equal(o, g, what=f"pml T2 {shape} after step(7)")
    o.update_fields(); g.update_fields()
    o.step(4); g.step(4)
    assert g.info().passes_t2 == 5
    assert_bit_equal(o, g, what=f"pml T2 {shape} after update_fields + step(4)")
    # the same solver without the two-step pass gives the same bits
    _, h = make_pair(Ni, Nj, Nk, d=(C, 1.25 * C, 0.8 * C), pml=pml, dtype=dtype, temporal=False)
    f = seeded_fields(91, (Nk, Nj, Ni), dtype=dtype, same_j=False)
    for c in range(9):
        h.upload(c, f[c])
    h.step(12)
    assert h.info().passes_t2 == 0
    for c in range(6):
        assert np.array_equal(h.download(c), g.download(c))


def test_pml_two_step_pass_device_source_across_shell_rim_core():
    """Device-resident source whose box straddles shell, rim and core cells, retiring in the middle of the run:
    the core takes the second step's J in the T2 kernel, the rim sweeps from the J arrays."""
    import math
    n, pml, steps, active = 48, 0.15, 11, 5
    o, g = make_pair(n, n, n, pml=pml)
    assert g.info().temporal == 1
    load_both(o, g, seeded_fields(17, (n, n, n)), comps=range(6))
    lo, hi = (4, 8, 9), (14, 12, 13)          # shell thickness 7: i = 4..6 shell, 7..9 rim, 10.. core
    w = [[0.3 + 0.1 * math.cos(0.7 * i) for i in range(lo[a], hi[a])] for a in range(3)]
    amp = [math.sin(0.4 * (t + 1)) for t in range(active)]
    g.set_source(lo, hi, w[0], w[1], w[2], amp)
    g.step(steps)
    for t in range(steps):
        for c in (6, 7, 8):
            j = o.field(c)
            j[...] = 0.0
            if t < active:
                for k in range(lo[2], hi[2]):
                    for jj in range(lo[1], hi[1]):
                        for i in range(lo[0], hi[0]):
                            j[k, jj, i] = ((amp[t] * w[0][i - lo[0]]) * w[1][jj - lo[1]]) * w[2][k - lo[2]]
        o.update_fields()
    assert_bit_equal(o, g, what="pml T2 device source")


def test_pml_explicit_thickness_matches_percent():
    Ni, Nj, Nk = 20, 20, 20
    o, g = make_pair(Ni, Nj, Nk, pml=0.2, pml_thickness=(4, 4, 4))
    load_both(o, g, seeded_fields(2, (Nk, Nj, Ni)))
    o.step(5)
    g.step(5)
    assert_bit_equal(o, g, what="pml thickness")


@pytest.mark.parametrize("name", ["random_periodic_16x12x10", "random_periodic_33x7x5", "random_pml_20x16x12", "random_pml_24x24x24"])
def test_golden_random_cases(name, golden_dir):
    """Committed outputs of the real reference (oracle/make_golden.py) -- no oracle in the loop."""
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    m = json.loads(str(z["meta"]))
    Ni, Nj, Nk = m["Ni"], m["Nj"], m["Nk"]
    p = fb.Parameters(Ni, Nj, Nk, 0, Ni * m["dx"], 0, Nj * m["dy"], 0, Nk * m["dz"], m["dx"], m["dy"], m["dz"])
    g = fb.FDTD(p, m["dt"], j_openmp_quirk=True) if m["pml_percent"] is None else fb.FDTD_PML(p, m["dt"], m["pml_percent"], j_openmp_quirk=True)
    f = seeded_fields(m["seed"], (Nk, Nj, Ni))
    for c in range(9):
        g.upload(c, f[c])
    done = 0
    names = ["EX", "EY", "EZ", "BX", "BY", "BZ"]
    for s in m["steps"]:
        g.step(s - done)
        done = s
        for c in range(6):
            assert np.array_equal(g.download(c), z[f"{names[c]}_step{s}"]), f"{name} {names[c]} step {s}"


@pytest.mark.parametrize("pml,name", [(None, "sample_32_100_periodic"), (0.2, "sample_32_100_pml02")])
@pytest.mark.parametrize("device_source", [False, True])
def test_sample_scenario_golden(pml, name, device_source, golden_dir):
    """perf-tests/sample/sample.cpp scenario, n=32, 100 steps, against the real reference's fields,
    with the host per-step J writes (scatter) and with the device-resident source."""
    n, steps = 32, 100
    z = np.load(os.path.join(golden_dir, name + "_fields.npz"))
    p = params(n, n, n)
    g = fb.FDTD(p, 0.2) if pml is None else fb.FDTD_PML(p, 0.2, pml)
    lo, hi, active, value = sample_source(n, steps)
    if device_source:
        import math
        PI, T, Tx = 3.14159265358, 8.0, 4.0 * C
        amp = [math.sin(2.0 * PI * (float(t + 1) * 0.2) / T) for t in range(active)]
        w = [[math.pow(math.cos(2.0 * PI * (float(i) * C) / Tx), 2.0) for i in range(lo[a], hi[a])] for a in range(3)]
        g.set_source(lo, hi, w[0], w[1], w[2], amp)
        g.step(steps)
    else:
        idx = np.array([i + j * n + k * n * n for k in range(lo[2], hi[2]) for j in range(lo[1], hi[1]) for i in range(lo[0], hi[0])])
        for t in range(active):
            vals = np.array([value(t, i % n, (i // n) % n, i // (n * n)) for i in idx])
            for c in (6, 7, 8):
                g.get_field(c)[idx] = vals
            g.update_fields()
        g.zeroed_currents()
        for t in range(active, steps):
            g.update_fields()
    for c, nm in enumerate(["EX", "EY", "EZ", "BX", "BY", "BZ"]):
        assert np.array_equal(g.download(c), z[nm]), f"{name} {nm} device_source={device_source}"
    # printed 10x10 slice of sample.cpp:125-134 (5 decimals) -- SURVEY.md B.3 row 6
    if pml is None:
        ex = g.download(0)
        row = [f"{v:.5f}" for v in ex[16, 16, 11:21]]
        assert row == ["0.00373", "-0.02977", "-0.02579", "-0.02357", "-0.02457", "-0.02457", "-0.02357", "-0.02579", "-0.02977", "0.00373"]


def test_convergence_unit_tests(golden_dir):
    """The reference's 12 Convergence.* tests (unit-tests/test_FDTD_method.cpp:71-203) on the GPU solver:
    ratio in 4.0 +- 0.1, and err_1/err_2 equal to the real reference's to the last bit."""
    import math
    gold = json.load(open(os.path.join(golden_dir, "convergence.json")))
    PI, T = 3.14159265358, 5e-13
    E_, B_ = {"EX": 0, "EY": 1, "EZ": 2}, {"BX": 3, "BY": 4, "BZ": 5}
    cases = {
        "x_axis_EY": (1, 5, 0, 1.0, 1, (16, 8, 4)), "x_axis_BZ": (1, 5, 0, 1.0, 5, (16, 8, 4)),
        "x_axis_EZ": (2, 4, 0, -1.0, 2, (16, 8, 4)), "x_axis_BY": (2, 4, 0, -1.0, 4, (16, 8, 4)),
        "y_axis_EX": (0, 5, 1, -1.0, 0, (8, 16, 4)), "y_axis_BZ": (0, 5, 1, -1.0, 5, (8, 16, 4)),
        "y_axis_EZ": (2, 3, 1, 1.0, 2, (8, 16, 4)), "y_axis_BX": (2, 3, 1, 1.0, 3, (8, 16, 4)),
        "z_axis_EX": (0, 4, 2, 1.0, 0, (4, 8, 16)), "z_axis_BY": (0, 4, 2, 1.0, 4, (4, 8, 16)),
        "z_axis_EY": (1, 3, 2, -1.0, 1, (4, 8, 16)), "z_axis_BX": (1, 3, 2, -1.0, 3, (4, 8, 16)),
    }

    def run(ef, bf, axis, sign, tf, N):
        Ni, Nj, Nk = N
        d = (1.0 / float(Ni), 2.0 / float(Nj), 3.0 / float(Nk))
        box = (0.0, 1.0, 0.0, 2.0, 0.0, 3.0)
        iters = 16 * (max(N) // 16)
        dt = T / float(iters)
        g = fb.FDTD(fb.Parameters(Ni, Nj, Nk, *box, *d), dt)
        a, b = box[2 * axis], box[2 * axis + 1]
        e = np.zeros((Nk, Nj, Ni)); bb = np.zeros((Nk, Nj, Ni))
        for m in range(N[axis]):        # Test_FDTD::initial_filling, src/FDTD/test_FDTD.cpp:5-51
            x = float(m) * d[axis]
            sl = [slice(None)] * 3
            sl[2 - axis] = m
            e[tuple(sl)] = sign * math.sin(2.0 * PI * (x - a) / (b - a))
            bb[tuple(sl)] = math.sin(2.0 * PI * (d[axis] / 2.0 + x - a) / (b - a))
        g.upload(ef, e); g.upload(bf, bb)
        for _ in range(iters):
            g.update_fields()
        f = g.download(tf)
        is_b = tf > 2                    # Test_FDTD::get_max_abs_error, src/FDTD/test_FDTD.cpp:89-130
        s, x, err = (1.0 if is_b else sign), (d[axis] / 2.0 if is_b else 0.0), 0.0
        for m in range(N[axis]):
            ix = [0, 0, 0]; ix[2 - axis] = m
            err = max(err, abs(s * f[tuple(ix)] - math.sin(2.0 * PI * (x - a - C * T) / (b - a))))
            x += d[axis]
        return err

    for name, (ef, bf, ax, sg, tf, N) in cases.items():
        e1 = run(ef, bf, ax, sg, tf, N)
        e2 = run(ef, bf, ax, sg, tf, tuple(2 * v for v in N))
        assert abs(e1 / e2 - 4.0) <= 0.1, name
        assert e1 == gold[name]["err1"] and e2 == gold[name]["err2"], name


def test_scatter_gather_and_zeroed_currents():
    Ni, Nj, Nk = 16, 8, 6
    o, g = make_pair(Ni, Nj, Nk)
    f = seeded_fields(9, (Nk, Nj, Ni), same_j=False)
    load_both(o, g, f)
    idx = np.array([0, 5, Ni * Nj * Nk - 1, 77, 300])
    for c in range(9):
        assert np.array_equal(g.get_field(c)[idx], f[c].reshape(-1)[idx])
    g.get_field(fb.Component.EZ)[idx] = np.arange(5, dtype=np.float64)
    o.field(2).reshape(-1)[idx] = np.arange(5, dtype=np.float64)
    o.step(2); g.step(2)
    assert_bit_equal(o, g, what="after scatter")
    o.zeroed_currents(); g.zeroed_currents()
    assert not g.download(6).any() and not g.download(7).any() and not g.download(8).any()
    o.step(3); g.step(3)
    assert_bit_equal(o, g, what="after zeroed_currents")
    assert g.get_field(0).size() == Ni * Nj * Nk


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_read_slice_and_dump(dtype, tmp_path):
    """fdtd_read_slice: device-side extraction of i / j / k slices equals slicing the dense download, after an odd
    number of steps (deferred B half step pending) and for J; dump_slices writes the OutFiles_<n>/<iter>.csv feed."""
    Ni, Nj, Nk = 24, 10, 7
    o, g = make_pair(Ni, Nj, Nk, dtype=dtype)
    f = seeded_fields(11, (Nk, Nj, Ni), dtype=dtype, same_j=False)
    load_both(o, g, f)
    o.step(3); g.step(3)
    for c in range(9):
        full = o.field(c)
        assert np.array_equal(g.read_slice(c, 2, 4), full[4]), c
        assert np.array_equal(g.read_slice(c, 1, 9), full[:, 9, :]), c
        assert np.array_equal(g.read_slice(c, 0, 23), full[:, :, 23]), c
    with pytest.raises(TypeError):
        g.read_slice(0, 2, Nk)
    paths = g.dump_slices(3, root=str(tmp_path))
    assert [os.path.relpath(p, tmp_path) for p in paths] == [f"OutFiles_{c}/3.csv" for c in range(1, 7)]
    rows = np.array([[float(v) for v in line.split(";")] for line in open(paths[4]).read().split()])
    assert np.array_equal(rows, o.field(4)[Nk // 2].astype(np.float64))
    o.step(1); g.step(1)
    assert_bit_equal(o, g, what="stepping after slice reads")


def test_errors_match_reference_semantics():
    with pytest.raises(ValueError, match="invalid parameters"):      # FDTD.cpp:5-7
        fb.FDTD(params(0, 4, 4), 0.2)
    with pytest.raises(ValueError, match="invalid parameters"):
        fb.FDTD(params(4, 4, 4), 0.0)
    g = fb.FDTD(params(4, 4, 4), 0.2)
    with pytest.raises(LookupError, match="Invalid field component"):  # FDTD.cpp:149
        g.get_field(9)
    with pytest.raises(TypeError):
        g.upload(0, np.zeros(5))


def test_large_grid_properties():
    """BASELINE size (256^3 fp64, config 1): size-independent properties instead of an oracle run --
    (a) fused and two-sweep paths agree bit for bit, (b) a uniform field is a fixed point (all curls
    vanish), (c) linearity: step(a*F) == a*step(F) for a power-of-two scale."""
    n, steps = 256, 6
    p = params(n, n, n)
    rng = np.random.default_rng(42)
    f = [rng.uniform(-1, 1, size=(n, n, n)) for _ in range(6)]
    outs = []
    for fusion, scale in ((True, 1.0), (False, 1.0), (True, 4.0)):
        g = fb.FDTD(p, 0.2, fusion=fusion)
        for c in range(6):
            g.upload(c, f[c] * scale)
        g.step(steps)
        outs.append([g.download(c) for c in range(6)])
        g.close()
    for c in range(6):
        assert np.array_equal(outs[0][c], outs[1][c]), "fused vs two-sweep"
        assert np.array_equal(outs[0][c] * 4.0, outs[2][c]), "linearity"
    g = fb.FDTD(p, 0.2)
    for c in range(6):
        g.upload(c, np.full((n, n, n), float(c + 1)))
    g.step(3)
    for c in range(6):
        assert np.array_equal(g.download(c), np.full((n, n, n), float(c + 1)))


# ---- temporally blocked two-step pass (fused_kernel_t2.cuh) ---------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape,steps", [((16, 12, 10), 8), ((64, 64, 64), 5), ((32, 8, 4), 9), ((2, 1, 4), 6), ((4, 2, 5), 7),
                                         ((60, 6, 5), 4), ((124, 20, 40), 6), ((8, 3, 70), 5), ((120, 9, 6), 2)])
def test_temporal_blocking_bit_exact(shape, steps, dtype):
    """fdtd_step(n) pairs steps into the T2 pass (two Yee steps per launch): must equal the oracle, and the
    one-step fused pass, bit for bit -- distinct static Jx/Jy/Jz, tiles that do not divide the grid, wraps."""
    Ni, Nj, Nk = shape
    o, g = make_pair(Ni, Nj, Nk, d=(C, 1.25 * C, 0.8 * C), dtype=dtype)
    _, g1 = make_pair(Ni, Nj, Nk, d=(C, 1.25 * C, 0.8 * C), dtype=dtype, temporal=False)
    f = seeded_fields(31, (Nk, Nj, Ni), dtype=dtype, same_j=False)
    load_both(o, g, f)
    for c in range(9):
        g1.upload(c, f[c])
    l0, l1 = g.info().launches, g1.info().launches
    o.step(steps); g.step(steps); g1.step(steps)
    if g.info().fused:   # (rows that are not a multiple of the vector width run the two-sweep kernels)
        assert g.info().launches - l0 == (steps + 1) // 2, "steps were not paired into T2 passes"
        assert g1.info().launches - l1 == steps
    assert_bit_equal(o, g, what=f"T2 {shape} {dtype}")
    assert_bit_equal(o, g1, what=f"T1 {shape} {dtype}")
    # a second batch starts from the pending-half-step state (n_half = 2 in stage A)
    o.step(3); g.step(3)
    assert_bit_equal(o, g, what=f"T2 second batch {shape}")


@pytest.mark.parametrize("variant", range(3))
def test_temporal_blocking_variants(variant, monkeypatch):
    monkeypatch.setenv("FDTD_B200_T2_VARIANT", str(variant))
    Ni, Nj, Nk = 68, 37, 21
    o, g = make_pair(Ni, Nj, Nk, d=(C, 1.25 * C, 0.8 * C))
    load_both(o, g, seeded_fields(17, (Nk, Nj, Ni), same_j=False))
    o.step(6); g.step(6)
    assert_bit_equal(o, g, what=f"T2 variant {variant}")


@pytest.mark.parametrize("active,steps", [(5, 12), (6, 12), (40, 100), (1, 4), (7, 7), (8, 7)])
@pytest.mark.parametrize("n", [32])
def test_temporal_blocking_device_source(n, active, steps):
    """Device-resident source under the T2 pass: the second step of a pair is injected in the kernel, pairs never
    straddle the step where the source retires, and the J arrays read back like the reference's."""
    import math
    PI, T, Tx = 3.14159265358, 8.0, 4.0 * C
    lo, hi, _, _ = sample_source(n, steps)
    amp = [math.sin(2.0 * PI * (float(t + 1) * 0.2) / T) for t in range(active)]
    w = [[math.pow(math.cos(2.0 * PI * (float(i) * C) / Tx), 2.0) for i in range(lo[a], hi[a])] for a in range(3)]
    o = Oracle(**sample_params(n))
    g = fb.FDTD(params(n, n, n), 0.2)
    rng = np.random.default_rng(5)
    for c in range(6):
        a = rng.uniform(-1, 1, size=(n, n, n))
        o.field(c)[...] = a
        g.upload(c, a)
    g.set_source(lo, hi, w[0], w[1], w[2], amp)
    for t in range(steps):
        if t < active:
            for k in range(lo[2], hi[2]):
                for j in range(lo[1], hi[1]):
                    for i in range(lo[0], hi[0]):
                        v = ((amp[t] * w[0][i - lo[0]]) * w[1][j - lo[1]]) * w[2][k - lo[2]]
                        for c in (6, 7, 8):
                            o.field(c)[k, j, i] = v
        elif t == active:
            o.zeroed_currents()
        o.update_fields()
    g.step(steps)
    assert_bit_equal(o, g, comps=range(9), what=f"T2 device source active={active} steps={steps}")


# ---- the benchmarked flavours of the T2 pass: TMA-fed tiles, both storage types (missing #3) -------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("currents", ["none", "box"])
@pytest.mark.parametrize("shape,steps", [((128, 48, 12), 6), ((192, 40, 9), 5), ((250, 30, 7), 4), ((252, 30, 7), 4),
                                         ((256, 64, 6), 4), ((136, 28, 20), 7)])
def test_temporal_blocking_tma_tiles(shape, steps, dtype, currents):
    """Grids wide enough that interior tiles take the TMA ring (footprint 64 cells x 16 rows without a periodic wrap:
    Ni >= 122, Nj >= 26) while edge tiles keep the per-thread cp.async ring -- fp64 (64-cell boxes) and fp32 (68-cell
    boxes starting 2 cells early).  `currents` = "none": every tile runs the J-free flavours; "box": J is non-zero on a
    small box only, so the tiles that meet it run the J flavour next to TMA tiles.  Ni = 250 / 252: the last tile column
    wraps (Ni is not a multiple of the 60-cell tile); fp32 needs Ni % 4 == 0 for the fused passes (250 -> sweep kernels)."""
    Ni, Nj, Nk = shape
    o, g = make_pair(Ni, Nj, Nk, d=(C, 1.25 * C, 0.8 * C), dtype=dtype)
    f = seeded_fields(57, (Nk, Nj, Ni), dtype=dtype, same_j=False)
    load_both(o, g, f, comps=range(6))
    if currents == "box":
        idx = np.array([i + j * Ni + k * Ni * Nj for k in (1, 2) for j in (20, 21, 22) for i in range(70, 76)])
        for c in (6, 7, 8):
            vals = f[c].reshape(-1)[idx]
            g.scatter(c, idx, vals)
            o.field(c).reshape(-1)[idx] = vals
    o.step(steps); g.step(steps)
    if g.info().fused:
        assert g.info().passes_t2 == steps // 2
    assert_bit_equal(o, g, what=f"T2 TMA tiles {shape} {np.dtype(dtype).name} J={currents}")
    o.step(3); g.step(3)
    assert_bit_equal(o, g, what=f"T2 TMA tiles, second batch {shape}")


def _checker(Ni, Nj, Nk, d, dt):
    """The real reference when oracle/_ref travelled to this box, else the C restatement in the same J semantics."""
    from oracle.pyoracle import Reference, have_reference
    if have_reference():
        return Reference(Ni, Nj, Nk, d[0], d[1], d[2], dt), "reference"
    return Oracle(Ni, Nj, Nk, d[0], d[1], d[2], dt, j_mode=J_OPENMP), "oracle"


@pytest.mark.parametrize("shape,steps", [((256, 256, 256), 10), ((512, 512, 64), 4)])
def test_baseline_size_against_reference(shape, steps):
    """BASELINE sizes against the reference itself (FDTD_openmp::FDTD, src/FDTD/FDTD.cpp:153-157, through oracle/_ref):
    256^3 x 10 steps (configs[1]) and a 64-plane slab of the 512^3 bench grid x 4 steps (configs[2]: random E/B seed 42,
    the sample's point source active every step), bit for bit."""
    Ni, Nj, Nk = shape
    d = (C, C, C)
    ref, kind = _checker(Ni, Nj, Nk, d, 0.2)
    g = fb.FDTD(params(Ni, Nj, Nk), 0.2, j_openmp_quirk=True)
    rng = np.random.default_rng(42)
    for c in range(6):
        a = rng.uniform(-1, 1, size=(Nk, Nj, Ni))
        ref.field(c)[...] = a
        g.upload(c, a)
    import math
    PI, T, Tx = 3.14159265358, 8.0, 4.0 * C
    lo = [N // 2 - 1 for N in (Ni, Nj, Nk)]
    hi = [N // 2 + 1 for N in (Ni, Nj, Nk)]
    w = [[math.pow(math.cos(2.0 * PI * (float(i - N // 2) * C) / Tx), 2.0) for i in range(l, h)] for l, h, N in zip(lo, hi, (Ni, Nj, Nk))]
    amp = [math.sin(2.0 * PI * (float(t + 1) * 0.2) / T) for t in range(steps)]
    g.set_source(lo, hi, w[0], w[1], w[2], amp)
    g.step(steps)
    for t in range(steps):
        for k in range(lo[2], hi[2]):
            for j in range(lo[1], hi[1]):
                for i in range(lo[0], hi[0]):
                    v = ((amp[t] * w[0][i - lo[0]]) * w[1][j - lo[1]]) * w[2][k - lo[2]]
                    for c in (6, 7, 8):
                        ref.field(c)[k, j, i] = v
        ref.update_fields()
    assert g.info().passes_t2 == steps // 2
    for c in range(6):
        assert np.array_equal(g.download(c), ref.field(c)), f"{shape} component {c} differs from the {kind}"
    ref.close()


def test_golden_kokkos_distinct_currents(golden_dir):
    """Committed outputs of the real FDTD_kokkos::FDTD with distinct Jx / Jy / Jz (kokkos_functors.h:81-89): the GPU
    solver's default J semantics, no oracle in the loop."""
    z = np.load(os.path.join(golden_dir, "random_periodic_kokkos_16x12x10.npz"))
    m = json.loads(str(z["meta"]))
    Ni, Nj, Nk = m["Ni"], m["Nj"], m["Nk"]
    for fusion in (True, False):
        g = fb.FDTD(fb.Parameters(Ni, Nj, Nk, 0, Ni * m["dx"], 0, Nj * m["dy"], 0, Nk * m["dz"], m["dx"], m["dy"], m["dz"]), m["dt"], fusion=fusion)
        f = seeded_fields(m["seed"], (Nk, Nj, Ni), same_j=False)
        for c in range(9):
            g.upload(c, f[c])
        done = 0
        for s in m["steps"]:
            g.step(s - done)
            done = s
            for c, nm in enumerate(["EX", "EY", "EZ", "BX", "BY", "BZ"]):
                assert np.array_equal(g.download(c), z[f"{nm}_step{s}"]), f"{nm} step {s} fusion={fusion}"
        g.close()


# ---- opt-in fp32 arithmetic (FDTD_FLAG_F32_ARITH): float storage AND float arithmetic, 4 cells per lane ---------------
@pytest.mark.parametrize("shape,steps,pml", [((16, 12, 10), 7, None), ((64, 64, 64), 5, None), ((128, 48, 12), 6, None), ((252, 30, 7), 4, None),
                                             ((256, 64, 6), 5, None), ((120, 9, 6), 3, None), ((36, 8, 4), 9, None),
                                             ((64, 48, 48), 6, 0.1), ((20, 16, 12), 5, 0.2)])
def test_f32_arith_mode_bit_exact_vs_float_oracle(shape, steps, pml):
    """Every kernel family of the mode (T2 pass with 128-cell TMA boxes, the two sweeps for single steps and the deferred
    half step, the PML rim sweeps) against the oracle's float-arithmetic restatement (same association, every operation
    rounded to float, no contraction): bit for bit, distinct Jx / Jy / Jz."""
    Ni, Nj, Nk = shape
    o, g = make_pair(Ni, Nj, Nk, d=(C, 1.25 * C, 0.8 * C), dtype=np.float32, pml=pml, f32_arith=True)
    assert g.info().f32_arith == 1
    load_both(o, g, seeded_fields(71, (Nk, Nj, Ni), dtype=np.float32, same_j=False))
    o.step(steps); g.step(steps)
    assert_bit_equal(o, g, what=f"f32 arithmetic {shape} pml={pml}")
    for _ in range(3):
        o.update_fields(); g.update_fields()
    assert_bit_equal(o, g, what=f"f32 arithmetic {shape}, update_fields loop")
    if pml is None and Ni % 4 == 0 and Nk >= 4:
        assert g.info().passes_t2 == steps // 2 + 1


@pytest.mark.parametrize("steps", [100, 1000])
def test_f32_arith_mode_within_north_star_tolerance(steps):
    """north_star's fp32 bar: <= 1e-5 relative L-inf against the fp64 reference semantics after N steps -- the sample
    scenario (perf-tests/sample/sample.cpp: point source for 40 steps, then free propagation), 100 and 1000 steps."""
    import math
    n = 32
    o = Oracle(**sample_params(n))
    run_sample(o, n, steps)
    g = fb.FDTD(params(n, n, n), 0.2, dtype=np.float32, f32_arith=True)
    lo, hi, active, _ = sample_source(n, steps)
    PI, T, Tx = 3.14159265358, 8.0, 4.0 * C
    amp = [math.sin(2.0 * PI * (float(t + 1) * 0.2) / T) for t in range(active)]
    w = [[math.pow(math.cos(2.0 * PI * (float(i) * C) / Tx), 2.0) for i in range(lo[a], hi[a])] for a in range(3)]
    g.set_source(lo, hi, w[0], w[1], w[2], amp)
    g.step(steps)
    worst = max(rel_linf(g.download(c), o.field(c)) for c in range(6))
    assert worst <= 1e-5, worst


def test_f32_arith_needs_float_storage():
    with pytest.raises(TypeError):
        fb.FDTD(params(8, 8, 8), 0.2, dtype=np.float64, f32_arith=True)


# ---- pending J writes: reference-style loops with a J write before every update_fields() still pair -----------------------
@pytest.mark.parametrize("shape,pml,dtype,f32_arith", [((32, 16, 12), None, np.float64, False), ((128, 48, 8), None, np.float64, False),
                                                        ((64, 24, 16), None, np.float32, False), ((64, 24, 16), None, np.float32, True),
                                                        ((64, 48, 48), 0.1, np.float64, False)])
def test_loop_with_J_writes_between_calls_pairs(shape, pml, dtype, f32_arith):
    """`J[idx] = v; update_fields();` repeated (perf-tests/sample/sample.cpp:66-87): a J write that arrives while one call is
    recorded becomes a pending box, the next call issues both steps as ONE two-step pass (stage A on the old J, stage B on the
    box), and every observable state equals the oracle's -- distinct Jx / Jy / Jz on top of a non-zero J background, writes
    that change every step, a read in the middle, a write outside the box (falls back), zeroed_currents at the end."""
    Ni, Nj, Nk = shape
    o, g = make_pair(Ni, Nj, Nk, d=(C, 1.25 * C, 0.8 * C), dtype=dtype, pml=pml, f32_arith=f32_arith)
    f = seeded_fields(13, (Nk, Nj, Ni), dtype=dtype, same_j=False)
    load_both(o, g, f)
    ci, cj, ck = Ni // 2, Nj // 2, Nk // 2
    idx = np.array([i + j * Ni + k * Ni * Nj for k in (ck - 1, ck) for j in (cj - 1, cj) for i in (ci - 1, ci)])
    rng = np.random.default_rng(5)
    p0 = g.info().passes_t2
    steps = 10
    for t in range(steps):
        for c in (6, 7, 8):
            v = rng.uniform(-1, 1, size=idx.size).astype(dtype)
            g.get_field(c)[idx] = v
            o.field(c).reshape(-1)[idx] = v
        g.update_fields(); o.update_fields()
    assert g.info().passes_t2 - p0 == steps // 2, "J writes between calls broke the pairing"
    assert_bit_equal(o, g, comps=range(9), what=f"J-write loop {shape}")
    # odd call + read in the middle: the recorded step runs alone on the old J, then the arrays take the pending writes
    for c in (6, 7, 8):
        g.get_field(c)[idx] = np.full(idx.size, 0.5, dtype=dtype); o.field(c).reshape(-1)[idx] = 0.5
    g.update_fields(); o.update_fields()
    for c in (6, 7, 8):
        g.get_field(c)[idx] = np.full(idx.size, -0.25, dtype=dtype); o.field(c).reshape(-1)[idx] = -0.25
    assert_bit_equal(o, g, comps=range(9), what="read with a recorded step and pending writes")
    # a write outside the box while writes are pending -> falls back, same results
    g.update_fields(); o.update_fields()
    far = np.array([1 + 2 * Ni + 1 * Ni * Nj])
    g.get_field(6)[idx] = np.full(idx.size, 0.125, dtype=dtype); o.field(6).reshape(-1)[idx] = 0.125
    g.get_field(7)[far] = np.array([2.0], dtype=dtype); o.field(7).reshape(-1)[far] = 2.0
    g.update_fields(); o.update_fields()
    g.update_fields(); o.update_fields()
    g.zeroed_currents(); o.zeroed_currents()
    g.step(3); o.step(3)
    assert_bit_equal(o, g, comps=range(9), what="after fallback, zeroed_currents and a batch")


def _random_cases(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for t in range(n):
        mode = ["f64", "f64", "f32", "f32a"][int(rng.integers(0, 4))]
        Ni = int(rng.integers(1, 90)) * (4 if mode != "f64" and rng.random() < 0.8 else 2 if rng.random() < 0.8 else 1)
        Ni = max(1, min(Ni, 260))
        Nj, Nk = int(rng.integers(1, 40)), int(rng.integers(1, 24))
        pml = [None, None, 0.1, 0.2][int(rng.integers(0, 4))]
        steps = int(rng.integers(1, 7))
        out.append((Ni, Nj, Nk, mode, pml, steps, int(rng.integers(0, 2 ** 31 - 1))))
    return out


# (FDTD_FUZZ_N scales both random sweeps for one-off hunts: profiles/pytest_fuzz_r02.log ran 300 + 300 cases)
_FUZZ_N = int(os.environ.get("FDTD_FUZZ_N", "0"))


@pytest.mark.parametrize("case", _random_cases(_FUZZ_N or 36, 20261018), ids=lambda c: f"{c[0]}x{c[1]}x{c[2]}-{c[3]}-pml{c[4]}-s{c[5]}")
def test_random_shapes_bit_exact(case):
    """Seeded random sweep over grid extents (degenerate, odd, wider than one tile), storage / arithmetic modes, PML
    percentages and step counts, a batch followed by single calls: every kernel family has to agree with the oracle."""
    Ni, Nj, Nk, mode, pml, steps, seed = case
    dtype = np.float64 if mode == "f64" else np.float32
    o, g = make_pair(Ni, Nj, Nk, d=(C, 1.25 * C, 0.8 * C), dtype=dtype, pml=pml, f32_arith=(mode == "f32a"))
    load_both(o, g, seeded_fields(seed % 1000, (Nk, Nj, Ni), dtype=dtype, same_j=False))
    o.step(steps); g.step(steps)
    assert_bit_equal(o, g, what=f"random {case} after step({steps})")
    for _ in range(3):
        o.update_fields(); g.update_fields()
    assert_bit_equal(o, g, what=f"random {case} after 3 single calls")


@pytest.mark.parametrize("seed", range(_FUZZ_N or 24))
def test_random_call_sequences_keep_the_state_machine_exact(seed):
    """API fuzz with a fixed seed: random interleavings of update_fields(), step(n), J writes (small boxes that stay
    pending, large ones that do not, single components), E / B writes, sparse and dense reads, zeroed_currents(), a device
    source -- after every read the GPU solver must equal the oracle driven by the same calls (lazy pairing, pending J box,
    deferred half step, J bounding box, lagging J arrays are all host-side state that only shows through the results)."""
    from tests.util import fuzz_call_sequence
    fuzz_call_sequence(seed)
