"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with the agreed keys,
uses every host thread even under torchrun's OMP_NUM_THREADS=1, and ranks other than 0 stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, args=()):
    env = dict(os.environ)
    env.update(extra_env or {})
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "32", "--steps", "2", "--warmup", "3", *args]
    return subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)


def test_reference_arm_json_line():
    r = _run({"OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Gcell-updates/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["steps"] == 2 and d["warmup"] == 3
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] == d["value"] and cb["sample"]
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if cb["kind"] == "reference":
        assert cb["cores"] == ncpu           # not the OMP_NUM_THREADS=1 torchrun exports
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0


def test_reference_arm_other_ranks_are_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, args=("--gpus", "2"))
    assert r.returncode == 0
    assert r.stdout.strip() == ""


def test_committed_bench_lines_carry_the_contract_keys():
    prof = os.path.join(ROOT, "profiles")
    seen = 0
    for name in sorted(os.listdir(prof)):
        if not (name.startswith("bench_r") and name.endswith(".json")):
            continue
        for line in open(os.path.join(prof, name)):
            if not line.startswith("{"):
                continue
            d = json.loads(line)
            seen += 1
            for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                      "dtype", "data", "config", "e2e", "gpu_launches"):
                assert k in d, (name, k)
            if d.get("impl") != "reference":
                assert d["gpu_launches"] > 0 and "roofline" in d and "clocks" in d, name
                for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
                    assert k in d["roofline"], (name, k)
    assert seen >= 4
