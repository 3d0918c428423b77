"""Where the end-to-end leg of bench.py spends its time (run on the GPU box).

    python tools/e2e_breakdown.py [--n 512] [--steps 50]
Times, with wall clock around synchronous C-ABI calls: upload / download of one field from pinned and from
pageable host memory, one scatter, one gather, update_fields() alone, and the per-step loop of bench.py.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd_method_b200 as fb  # noqa: E402

C = 3e10


def wall(fn, reps=1):
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--steps", type=int, default=50)
    a = ap.parse_args()
    n = a.n
    p = fb.Parameters(n, n, n, 0, n * C, 0, n * C, 0, n * C, C, C, C)
    g = fb.FDTD(p, 0.2)
    pinned = torch.empty((n, n, n), dtype=torch.float64).pin_memory()
    pinned.numpy()[...] = 0.5
    pageable = np.full((n, n, n), 0.25)
    gb = pinned.numpy().nbytes / 1e9
    out = {"n": n, "field_GB": gb}
    g.upload(0, pinned.numpy())
    out["upload_pinned_GBs"] = gb / wall(lambda: g.upload(0, pinned.numpy()), 3)
    out["download_pinned_GBs"] = gb / wall(lambda: g.download(0, pinned.numpy()), 3)
    out["upload_pageable_GBs"] = gb / wall(lambda: g.upload(1, pageable), 2)
    out["download_pageable_GBs"] = gb / wall(lambda: g.download(1, pageable), 2)
    # raw torch copy for comparison (same pinned buffer, contiguous cudaMemcpyAsync)
    dev = torch.empty((n, n, n), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()

    def tcopy_h2d():
        dev.copy_(pinned, non_blocking=True)
        torch.cuda.synchronize()

    def tcopy_d2h():
        pinned.copy_(dev, non_blocking=True)
        torch.cuda.synchronize()

    tcopy_h2d()
    out["torch_h2d_pinned_GBs"] = gb / wall(tcopy_h2d, 3)
    out["torch_d2h_pinned_GBs"] = gb / wall(tcopy_d2h, 3)
    del dev

    idx = np.arange(8, dtype=np.int64) + n * n * (n // 2) + n * (n // 2) + n // 2
    vals = np.ones(8)
    g.scatter(6, idx, vals)
    out["scatter_us"] = 1e6 * wall(lambda: g.scatter(6, idx, vals), 50)
    pidx = np.arange(100, dtype=np.int64) + n * n * (n // 2)
    g.gather(0, pidx)
    out["gather_us"] = 1e6 * wall(lambda: g.gather(0, pidx), 50)

    def one_step():
        g.update_fields()
        g.sync()

    one_step()
    out["update_fields_plus_sync_ms"] = 1e3 * wall(one_step, 10)

    def loop_step():
        for c in (6, 7, 8):
            g.scatter(c, idx, vals)
        g.update_fields()
        g.gather(0, pidx)

    loop_step()
    out["bench_loop_step_ms"] = 1e3 * wall(loop_step, a.steps)

    def block():
        g.step(a.steps)
        g.sync()

    block()
    out["step_block_ms_per_step"] = 1e3 * wall(block, 1) / a.steps
    def loop():
        for _ in range(a.steps):
            g.update_fields()      # reference-style caller loop: pairs lazily into the two-step pass
        g.sync()

    loop()
    out["update_fields_loop_ms_per_step"] = 1e3 * wall(loop, 1) / a.steps
    print(json.dumps(out), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/e2e_breakdown.json", "w") as fh:
        json.dump(out, fh)
    g.close()
