"""CPU-side checks of the C ABI: the library loads, exports every symbol include/fdtd_b200.h declares, the
host-only helpers agree with the oracle, and compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle.pyoracle import C, Oracle

import fdtd_method_b200 as fb
from fdtd_method_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "fdtd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fdtd_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _capi.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/fdtd_b200.h but not exported"
    assert set(names) == set(_capi.SIGNATURES), "python binding table out of sync with the header"
    assert L.fdtd_version() == 100


def test_params_layout_matches_reference_struct():
    # FDTD_struct::Parameters with FP=double: 3 ints, 4 bytes padding, 9 doubles
    assert ctypes.sizeof(fb.Parameters) == 88
    assert fb.Parameters.dx.offset == 64 and fb.Parameters.ax.offset == 16
    assert [int(c) for c in fb.Component] == list(range(9))


@pytest.mark.parametrize("N,pct,d", [(32, 0.2, C), (100, 0.07, 2.5 * C), (16, 0.0, C), (9, 0.5, C), (512, 0.0625, C)])
def test_pml_profile_matches_oracle(N, pct, d):
    p = _capi.lib().fdtd_pml_thickness(N, pct)
    o = Oracle(N, N, N, d, d, d, 0.2, pml_percent=pct)
    assert p == o.pml_size(0) == int(float(N) * pct)
    if 2 * p > N:
        return
    sigma, decay, coef2 = fb.pml_profile(N, p, d, 0.2)
    so, do, co = o.pml_tables(0)
    assert np.array_equal(sigma, so) and np.array_equal(decay, do) and np.array_equal(coef2, co)


def test_slab_ranges_tile_the_grid():
    for Nk in (1, 7, 512, 1000):
        for P in (1, 2, 3, 8):
            if P > Nk:
                continue
            r = [fb.slab_range(Nk, k, P) for k in range(P)]
            assert r[0][0] == 0 and r[-1][1] == Nk
            assert all(r[i][1] == r[i + 1][0] for i in range(P - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1


def test_invalid_parameters_raise_like_the_reference():
    with pytest.raises(ValueError, match="ERROR: invalid parameters"):
        fb.FDTD(fb.Parameters(0, 4, 4, 0, 1, 0, 1, 0, 1, 1, 1, 1), 0.1)
    with pytest.raises(ValueError, match="ERROR: invalid parameters"):
        fb.FDTD(fb.Parameters(4, 4, 4, 0, 1, 0, 1, 0, 1, 1, 1, 1), -1.0)


def test_no_cpu_fallback():
    """Without a GPU every solver construction must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fb.FDTD(fb.Parameters(4, 4, 4, 0, 1, 0, 1, 0, 1, 1, 1, 1), 0.1)


def test_product_never_imports_the_oracle():
    """The package and bench/product paths must not reference oracle/ (checker only)."""
    pkg = os.path.join(ROOT, "fdtd_method_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "fdtd_oracle" not in txt, f


def test_cost_weighted_slabs_for_pml():
    """fdtd_slab_range_cfg: uniform unless the solver has a PML shell in k; then the ranks that own k-shell planes get
    fewer planes (36 vs 6 words per cell-step), the ranges still tile the grid and every rank keeps >= 4 planes."""
    import ctypes
    L = _capi.lib()

    def ranges(Ni, Nj, Nk, n, mode, thick=(0, 0, 0), pct=0.0, flags=0):
        cfg = _capi.Config()
        L.fdtd_config_init(ctypes.byref(cfg))
        cfg.grid = fb.Parameters(Ni, Nj, Nk, 0, 1, 0, 1, 0, 1, 1, 1, 1)
        cfg.dt, cfg.nranks, cfg.pml_mode, cfg.pml_percent, cfg.flags = 0.2, n, mode, pct, flags
        for a in range(3):
            cfg.pml_thickness[a] = thick[a]
        out = []
        for r in range(n):
            b, e = ctypes.c_int(), ctypes.c_int()
            L.fdtd_slab_range_cfg(ctypes.byref(cfg), r, ctypes.byref(b), ctypes.byref(e))
            out.append((b.value, e.value))
        return out

    for n in (1, 2, 3, 4, 8):
        Nk = 512 * n
        assert ranges(512, 512, Nk, n, _capi.PML_NONE) == [fb.slab_range(Nk, r, n) for r in range(n)]
        assert ranges(512, 512, Nk, n, _capi.PML_THICKNESS, (32, 32, 32), flags=_capi.FLAG_UNIFORM_SLABS) == [fb.slab_range(Nk, r, n) for r in range(n)]
        w = ranges(512, 512, Nk, n, _capi.PML_THICKNESS, (32, 32, 32))
        assert w[0][0] == 0 and w[-1][1] == Nk and all(a[1] == b[0] for a, b in zip(w, w[1:]))
        assert all(e - b >= 4 for b, e in w)
        if n >= 3:
            h = [e - b for b, e in w]
            assert h[0] < h[1] and h[-1] < h[-2] and h[0] == h[-1]        # the k-shell owners are shorter
            # cost balance within 1.5 % (shell cell 6, core cell 1)
            core = 448 * 448
            main, shell = (512 * 512 - core) * 6 + core, 512 * 512 * 6
            cost = [sum(shell if (k < 32 or k >= Nk - 32) else main for k in range(b, e)) for b, e in w]
            assert max(cost) / min(cost) < 1.015, cost
    # no shell in k -> uniform; tiny grids -> uniform
    assert ranges(64, 64, 64, 4, _capi.PML_THICKNESS, (8, 8, 0)) == [fb.slab_range(64, r, 4) for r in range(4)]
    assert ranges(16, 16, 16, 4, _capi.PML_PERCENT, pct=0.2) == [fb.slab_range(16, r, 4) for r in range(4)]


def test_cpp_callers_compile_against_the_headers():
    """The drop-in C++ classes are header-only: every caller program under cpp/ (sample clone, kokkos_sample clone,
    convergence unit tests, coarray-style program) must compile and link against the in-tree library without a GPU."""
    import shutil
    import subprocess
    if shutil.which("g++") is None or shutil.which("make") is None:
        pytest.skip("no g++ / make")
    _capi.lib()   # builds libfdtd_b200.so if it is missing
    r = subprocess.run(["make", "-B", "-C", os.path.join(ROOT, "cpp")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    for exe in ("test_FDTD_method_b200", "sample_b200", "kokkos_sample_b200", "fdtd_coarray_b200"):
        assert os.path.exists(os.path.join(ROOT, "cpp", "bin", exe)), exe
    # without a GPU the programs fail loudly (no CPU fallback) instead of computing anything
    r = subprocess.run([os.path.join(ROOT, "cpp", "bin", "sample_b200")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        assert r.returncode != 0 and "no CUDA device" in r.stdout, r.stdout[-500:]


def test_t2_chunk_plan_tiles_every_plane_once():
    """The plane-chunk list of a two-step-pass launch (csrc/fdtd_capi.cu::t2_chunk_plan, host-only): every plane of the
    requested range(s) is produced by exactly one chunk, the list fits the kernel's table, and when the kernel waits for the
    halo itself only the thin chunks at the END of the list reach ghost planes."""
    import ctypes
    L = _capi.lib()
    lo_a, hi_a = (ctypes.c_int * 72)(), (ctypes.c_int * 72)()

    def plan(nk, lo, hi, lo2=0, hi2=0, wait=0, tiles=387, gx=9, kc=0):
        n = L.fdtd_debug_t2_chunk_plan(nk, lo, hi, lo2, hi2, wait, tiles, gx, kc, lo_a, hi_a, 72)
        assert 0 < n <= 72, (nk, lo, hi, n)
        return [(lo_a[i], hi_a[i]) for i in range(n)]

    rng = np.random.default_rng(1)
    cases = [(512, 0, 512, 0, 0), (1024, 0, 1024, 0, 0), (128, 0, 128, 0, 0), (16, 0, 16, 0, 0), (4, 0, 4, 0, 0), (5, 0, 5, 0, 0),
             (512, 2, 510, 0, 0), (512, 0, 2, 510, 512), (512, 34, 478, 0, 0), (470, 34, 470, 0, 0), (7, 0, 7, 0, 0)]
    for _ in range(200):
        nk = int(rng.integers(4, 1100))
        lo = int(rng.integers(0, nk - 1)); hi = int(rng.integers(lo + 1, nk + 1))
        cases.append((nk, lo, hi, 0, 0))
    for nk, lo, hi, lo2, hi2 in cases:
        for wait in (0, 1):
            for tiles, gx in ((387, 9), (1548, 18), (16, 4), (215, 5)):
                for kc in (0, 7, 64):
                    ch = plan(nk, lo, hi, lo2, hi2, wait, tiles, gx, kc)
                    want = sorted(list(range(lo, hi)) + list(range(lo2, hi2)))
                    got = sorted(k for a, b in ch for k in range(a, b))
                    assert got == want, (nk, lo, hi, lo2, hi2, wait, ch)
                    assert all(b > a for a, b in ch)
                    if wait and hi2 <= lo2 and hi - lo >= 16:
                        touching = [i for i, (a, b) in enumerate(ch) if a - 2 < 0 or b + 1 >= nk]
                        n_expected = (1 if lo - 2 < 0 else 0) + (1 if hi + 1 >= nk else 0)
                        assert touching == list(range(len(ch) - n_expected, len(ch))), (nk, lo, hi, ch)
                        assert all(ch[i][1] - ch[i][0] == 2 for i in touching)
    # the benchmarked shapes: 512^3 on one GPU = 3 chunks of 171; a 512-plane slab rank = interior chunks + 2 thin ones
    assert [b - a for a, b in plan(512, 0, 512)] == [171, 171, 170]
    p8 = plan(512, 0, 512, wait=1)
    assert p8[-2:] == [(510, 512), (0, 2)] and p8[0][0] == 2 and p8[-3][1] == 510


def test_one_process_ring_issues_collective_work_on_every_slab_first():
    """FDTDMulti._prep (no GPU: mock slabs): before a call that may wait for one slab, every slab issues its recorded work;
    accesses that make the library apply the deferred B half step -- any B access and WRITES of E -- flush on every slab
    first (the half step needs the ring exchange; issued on one slab and waited for at once it deadlocks)."""
    from fdtd_method_b200.multi import FDTDMulti

    class Slab:
        def __init__(self):
            self.calls = []
        def issue(self):
            self.calls.append("issue")
        def flush(self):
            self.calls.append("flush")

    m = FDTDMulti.__new__(FDTDMulti)
    m.slabs = [Slab(), Slab(), Slab()]
    expect = {(0, False): "issue", (0, True): "flush", (2, True): "flush", (3, False): "flush", (5, True): "flush",
              (6, True): "issue", (8, False): "issue", (None, False): "issue"}
    for (comp, write), want in expect.items():
        for s in m.slabs:
            s.calls.clear()
        m._prep(comp, write=write)
        assert all(s.calls == [want] for s in m.slabs), (comp, write, [s.calls for s in m.slabs])
