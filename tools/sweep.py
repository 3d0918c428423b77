"""Kernel-variant sweep (run on the GPU box): times the fused pass variants and the two-sweep path.

    python tools/sweep.py [--n 512] [--dtype f64] [--steps 20]
Writes one JSON line per configuration to stdout and gpurun_out/sweep.jsonl.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd_method_b200 as fb  # noqa: E402

C = 3e10


def time_config(n, dtype, fusion, steps, warmup=5, pml=None):
    p = fb.Parameters(n, n, n, 0, n * C, 0, n * C, 0, n * C, C, C, C)
    g = fb.FDTD(p, 0.2, dtype=dtype, fusion=fusion) if pml is None else fb.FDTD_PML(p, 0.2, pml, dtype=dtype)
    rng = np.random.default_rng(0)
    plane = rng.uniform(-1, 1, size=(n, n, n)).astype(dtype)
    for c in range(6):
        g.upload(c, plane)
    g.step(warmup)
    g.sync()
    g.timer_start()
    g.step(steps)
    ms = g.timer_stop()
    g.close()
    return ms / steps


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="+", default=[512])
    ap.add_argument("--dtype", nargs="+", default=["f64"])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--variants", type=int, nargs="+", default=[0, 1, 2, 3, 4, 5])
    ap.add_argument("--kc", type=int, nargs="+", default=[0])
    ap.add_argument("--pml", action="store_true")
    ap.add_argument("--t2", type=int, nargs="*", default=[], help="T2 (two-step pass) variants to time")
    ap.add_argument("--no-sweeps", action="store_true", help="skip the two-sweep and one-step rows")
    a = ap.parse_args()
    os.makedirs("gpurun_out", exist_ok=True)
    out = open("gpurun_out/sweep.jsonl", "a")
    for n in a.n:
        for dt in a.dtype:
            dtype = np.float64 if dt == "f64" else np.float32
            W = 8 if dt == "f64" else 4
            rows = []
            os.environ["FDTD_B200_NO_T2"] = "1"
            if not a.no_sweeps:
                ms = time_config(n, dtype, False, a.steps)
                rows.append(dict(n=n, dtype=dt, path="two-sweep", ms=ms))
            for v in ([] if a.no_sweeps else a.variants):
                for kc in a.kc:
                    os.environ["FDTD_B200_FUSED_VARIANT"] = str(v)
                    os.environ["FDTD_B200_FUSED_KC"] = str(kc)
                    ms = time_config(n, dtype, True, a.steps)
                    rows.append(dict(n=n, dtype=dt, path=f"fused v{v} kc{kc}", ms=ms))
            os.environ["FDTD_B200_NO_T2"] = "0"
            for v in a.t2:
                for kc in a.kc:
                    os.environ["FDTD_B200_T2_VARIANT"] = str(v)
                    os.environ["FDTD_B200_FUSED_KC"] = str(kc)
                    ms = time_config(n, dtype, True, a.steps)
                    rows.append(dict(n=n, dtype=dt, path=f"T2 v{v} kc{kc}", ms=ms))
            if a.pml:
                ms = time_config(n, dtype, False, max(3, a.steps // 4), pml=0.0625)
                rows.append(dict(n=n, dtype=dt, path="pml 0.0625 two-sweep", ms=ms))
            for r in rows:
                r["gcells"] = n ** 3 / r["ms"] / 1e6
                r["alg_GBs_21w"] = r["gcells"] * 21 * W
                r["moved_GBs_12w"] = r["gcells"] * 12 * W
                line = json.dumps(r)
                print(line, flush=True)
                out.write(line + "\n")
    out.close()
