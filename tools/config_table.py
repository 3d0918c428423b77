"""One-GPU table of the BASELINE.json configurations (run on the GPU box).

    python tools/config_table.py [--quick] > gpurun_out/config_table.jsonl

Rows: C2 256^3 fp64 periodic (1000 steps), C3 512^3 fp64 / fp32 with the sample source active, C4's single-GPU
point 1024^3 fp64 (zero fields + source: 8 GiB-per-array host uploads are skipped), C5's
single-GPU point 512^3 fp64 with 32-cell PML.  Device-resident inputs, CUDA events around fdtd_step(n).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fdtd_method_b200 as fb  # noqa: E402
from bench import C, sample_source_tables  # noqa: E402


def run(name, n, dtype, steps, warmup=10, pml=None, random_init=True, source=True, reps=3, f32_arith=False):
    p = fb.Parameters(n, n, n, -n / 2 * C, n / 2 * C, -n / 2 * C, n / 2 * C, -n / 2 * C, n / 2 * C, C, C, C)
    npdt = np.float64 if dtype == "f64" else np.float32
    kw = dict(dtype=npdt, f32_arith=f32_arith)
    g = fb.FDTD(p, 0.2, **kw) if pml is None else fb.FDTD_PML(p, 0.2, pml_thickness=(pml, pml, pml), **kw)
    if random_init:
        rng = np.random.default_rng(42)
        a = rng.uniform(-1, 1, size=(n, n, n)).astype(npdt)
        for c in range(6):
            g.upload(c, a)
            a = np.roll(a, 7 + c)
    if source:
        lo, hi, w, amp = sample_source_tables((n, n, n), warmup + reps * steps + 2)
        g.set_source(lo, hi, w[0], w[1], w[2], amp)
    g.step(warmup)
    g.sync()
    ms = []
    for _ in range(reps):
        g.timer_start()
        g.step(steps)
        g.flush()            # the closing B half step is inside the timed region
        ms.append(g.timer_stop() / steps)
    info = g.info()
    g.close()
    med = float(np.median(ms))
    W = 8 if dtype == "f64" else 4
    if pml:
        npml = n ** 3 - (n - 2 * pml) ** 3
        alg = ((n ** 3 - npml) * 21 + npml * 36) * W
    else:
        alg = n ** 3 * 21 * W
    row = dict(config=name, n=n, dtype=dtype, pml=pml, steps=steps, reps=reps, ms_per_step=med, ms_all=ms,
               gcells=n ** 3 / med / 1e6, alg_GBs=alg / med / 1e6, fused=int(info.fused), passes_t2=int(info.passes_t2),
               device_gib=info.device_bytes / 2 ** 30)
    print(json.dumps(row), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--skip-1024", action="store_true")
    a = ap.parse_args()
    q = a.quick
    run("C2 256^3 f64 periodic", 256, "f64", 200 if q else 1000, source=False)
    run("C3 512^3 f64 source", 512, "f64", 50 if q else 200)
    run("C3 512^3 f32 source", 512, "f32", 50 if q else 200)
    run("C3 512^3 f32 source, FDTD_FLAG_F32_ARITH (opt-in float arithmetic, not a reference mode)", 512, "f32", 50 if q else 200, f32_arith=True)
    run("C5/P=1 512^3 f64 PML 32", 512, "f64", 20 if q else 100, pml=32)
    run("C5/P=1 512^3 f32 PML 32", 512, "f32", 20 if q else 100, pml=32)
    run("C5/P=1 512^3 f32 PML 32, FDTD_FLAG_F32_ARITH", 512, "f32", 20 if q else 100, pml=32, f32_arith=True)
    if not a.skip_1024:
        run("C4/P=1 1024^3 f64 periodic (zero init + source)", 1024, "f64", 10 if q else 40, warmup=4, random_init=False)
