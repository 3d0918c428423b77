"""The C++ drop-in classes (include/FDTD_b200/*.h over the C ABI) driven by the reference's own callers:
the 12 convergence unit tests and the perf-tests/sample scenario, compiled with g++ on the box."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cpp_bins():
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "cpp")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return os.path.join(ROOT, "cpp", "bin")


def test_convergence_unit_tests_cpp(cpp_bins, golden_dir):
    r = subprocess.run([os.path.join(cpp_bins, "test_FDTD_method_b200")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    assert "15 tests, 0 failed" in r.stdout
    # err_1 / err_2 printed with 17 digits must equal the real reference's (tests/golden/convergence.json)
    gold = json.load(open(os.path.join(golden_dir, "convergence.json")))
    lines = r.stdout.splitlines()
    for name, g in gold.items():
        i = next(k for k, l in enumerate(lines) if l.startswith("[ RUN") and l.endswith("Convergence_b200." + name))
        err = next(l for l in lines[i:] if "err_1" in l).split()
        assert float(err[2]) == g["err1"] and float(err[5]) == g["err2"], name


def _slice_rows(text, after=None):
    lines = text.splitlines()
    if after is not None:
        lines = lines[next(k for k, l in enumerate(lines) if l.startswith(after)) + 1:]
    rows = [l.split() for l in lines if len(l.split()) == 10 and all(c in "-.0123456789" for c in "".join(l.split()))]
    return rows[:10]


def test_sample_clone_prints_the_reference_slice(cpp_bins, golden_dir):
    """./sample_b200 (default 32, 100): stdout format of perf-tests/sample/sample.cpp and the values of the real
    reference (tests/golden/sample_32_100_*.json holds its Ex slices), periodic and PML."""
    r = subprocess.run([os.path.join(cpp_bins, "sample_b200"), "pml"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    assert r.stdout.startswith("Execution time: ")
    assert "Execution time (PML): " in r.stdout
    per = json.load(open(os.path.join(golden_dir, "sample_32_100_periodic.json")))["slice_EX_xy"]
    pml = json.load(open(os.path.join(golden_dir, "sample_32_100_pml02.json")))["slice_EX_xy"]
    got = _slice_rows(r.stdout)
    assert got == [[f"{v:.5f}" for v in row] for row in per]
    got_pml = _slice_rows(r.stdout, after="PML:")
    assert got_pml == [[f"{v:.5f}" for v in row] for row in pml]
    # first row of ./sample (32,100)
    assert got[0] == ["0.00416", "0.01664", "0.00905", "0.01396", "0.00923", "0.01133", "0.01142", "0.02038", "0.00746", "0.01847"]


def test_sample_clone_kokkos_slice(cpp_bins, golden_dir):
    r = subprocess.run([os.path.join(cpp_bins, "sample_b200"), "32", "100", "--kokkos-slice"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    yz = json.load(open(os.path.join(golden_dir, "sample_32_100_periodic.json")))["slice_EX_yz"]
    assert _slice_rows(r.stdout) == [[f"{v:.5f}" for v in row] for row in yz]
    # first row of ./kokkos_sample (32,100)
    assert _slice_rows(r.stdout)[0][:3] == ["0.04722", "0.00665", "-0.01068"]


def test_sample_clone_writes_the_visualisation_csvs(cpp_bins, golden_dir, tmp_path):
    """`./sample N iters 1` (the argv python_script_legend/visualization.py:52 passes) writes
    OutFiles_<1..6>/<iter>.csv, ';'-separated k = N/2 slices; the last Ex file holds the golden slice values."""
    r = subprocess.run([os.path.join(cpp_bins, "sample_b200"), "32", "100", "1"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=300, cwd=tmp_path)
    assert r.returncode == 0, r.stdout
    for c in range(1, 7):
        files = sorted(os.listdir(tmp_path / f"OutFiles_{c}"), key=lambda f: int(f.split(".")[0]))
        assert files == [f"{t}.csv" for t in range(100)]
    rows = [[float(v) for v in line.split(";")] for line in open(tmp_path / "OutFiles_1" / "99.csv").read().split()]
    assert len(rows) == 32 and all(len(row) == 32 for row in rows)
    per = json.load(open(os.path.join(golden_dir, "sample_32_100_periodic.json")))["slice_EX_xy"]
    got = [[f"{rows[j][i]:.5f}" for i in range(11, 21)] for j in range(11, 21)]
    assert got == [[f"{v:.5f}" for v in row] for row in per]
    # the slice printed on stdout is the same data
    assert _slice_rows(r.stdout) == got


def test_kokkos_sample_clone(cpp_bins, golden_dir):
    """./kokkos_sample_b200: handles held across steps, operator(), per-step re-zero after the source window
    (kokkos_sample.cpp:82-84,105-107,114-129) -- prints the YZ slice of the real reference."""
    r = subprocess.run([os.path.join(cpp_bins, "kokkos_sample_b200")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    assert r.stdout.startswith("Execution time: ")
    yz = json.load(open(os.path.join(golden_dir, "sample_32_100_periodic.json")))["slice_EX_yz"]
    assert _slice_rows(r.stdout) == [[f"{v:.5f}" for v in row] for row in yz]
    assert _slice_rows(r.stdout)[0][:3] == ["0.04722", "0.00665", "-0.01068"]


@pytest.mark.parametrize("ngpu", [2, 4])
@pytest.mark.parametrize("prog", ["sample_b200", "kokkos_sample_b200"])
def test_drop_in_class_on_several_gpus(cpp_bins, golden_dir, gpu_count, ngpu, prog):
    """FDTD_b200::FDTD / FDTD_PML spanning several GPUs in ONE process (FDTD_B200_GPUS): the unmodified caller programs
    print the real reference's slices -- z slabs, copy-engine halo ring, get_field gathering the slabs."""
    if gpu_count < ngpu:
        pytest.skip(f"needs {ngpu} GPUs")
    args = ["pml"] if prog == "sample_b200" else []
    r = subprocess.run([os.path.join(cpp_bins, prog), *args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300,
                       env=dict(os.environ, FDTD_B200_GPUS=str(ngpu)))
    assert r.returncode == 0, r.stdout
    gold = json.load(open(os.path.join(golden_dir, "sample_32_100_periodic.json")))
    if prog == "sample_b200":
        pml = json.load(open(os.path.join(golden_dir, "sample_32_100_pml02.json")))["slice_EX_xy"]
        assert _slice_rows(r.stdout) == [[f"{v:.5f}" for v in row] for row in gold["slice_EX_xy"]]
        assert _slice_rows(r.stdout, after="PML:") == [[f"{v:.5f}" for v in row] for row in pml]
    else:
        assert _slice_rows(r.stdout) == [[f"{v:.5f}" for v in row] for row in gold["slice_EX_yz"]]


@pytest.mark.parametrize("images", [1, 2, 4])
def test_coarray_program(cpp_bins, golden_dir, gpu_count, images):
    """coarray/fdtd.F90's program (one image per GPU): banner lines and the 10x10 Ex slice, which for
    the 32^3 x 100 scenario is the slice the real reference's sample prints (same source, same steps)."""
    if gpu_count < images:
        pytest.skip(f"needs {images} GPUs")
    r = subprocess.run([os.path.join(cpp_bins, "fdtd_coarray_b200"), "--images", str(images), "32", "32", "32", "100"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    lines = r.stdout.splitlines()
    assert f"Running on {images} images" in lines
    assert any(l.startswith("Total execution time: ") and l.endswith(" seconds") for l in lines)
    per = json.load(open(os.path.join(golden_dir, "sample_32_100_periodic.json")))["slice_EX_xy"]
    assert _slice_rows(r.stdout) == [[f"{v:.5f}" for v in row] for row in per]
