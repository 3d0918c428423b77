#!/usr/bin/env python
"""bench.py -- Gcell-updates/s of the Yee time step (update_fields) on B200, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 512] [--dtype f64|f32]
                    [--workload periodic|pml]

One "step" = one update_fields() (B half, E, B half) over the whole grid.  Workload at N=1: the
configuration BASELINE.json's metric is quoted on, a 512^3 fp64 Yee grid, periodic boundaries, random
initial E/B (seed 42) plus the reference sample's point current source kept active (configs[2]); at N>1
the same 512^3 block PER GPU, z-slab partitioned (weak scaling, 512 x 512 x 512*N), one process per GPU,
halo planes over NCCL send/recv.  One JSON line on stdout (rank 0).

`--impl reference` times the reference's own CPU implementation (oracle/_ref/libfdtd_ref.so = the unmodified
src/FDTD/FDTD.cpp compiled by oracle/Makefile; the C oracle port when that file is absent) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C = 3e10
PI = 3.14159265358
WORDS_PER_CELL_STEP = 21          # E sweep 12 words + B sweep 9 words, periodic/interior cell
WORDS_PER_PML_CELL_STEP = 36
FALLBACK_HBM_GBS = 6650.0         # /opt/skills/guides/B200_PROFILING.md fallback


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(dtype, n):
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu --set full capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            d = json.load(fh)
        return d.get(f"{dtype}_{n}", {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def sample_source_tables(n_global, nsteps):
    """Source of perf-tests/sample/sample.cpp:15-31,57-63 as the tables fdtd_set_source
    takes, kept active for `nsteps` steps (configs[2]: 'with current-source injection')."""
    Ni, Nj, Nk = n_global
    T, dt, Tp, d = 8.0, 0.2, 4.0 * C, C
    lo, hi, w = [], [], []
    for N in (Ni, Nj, Nk):
        a = -(N / 2.0) * d
        l, h = int(math.floor((-Tp / 4.0 - a) / d)), int(math.floor((Tp / 4.0 - a) / d))
        lo.append(l), hi.append(h)
        w.append([math.pow(math.cos(2.0 * PI * (float(i) * d) / Tp), 2.0) for i in range(l, h)])
    amp = [math.sin(2.0 * PI * (float(t + 1) * dt) / T) for t in range(nsteps)]
    return lo, hi, w, amp


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------------
# CPU side: the reference's own implementation on the host cores
# --------------------------------------------------------------------------------------------------------
def cpu_solver(shape, pml):
    """(solver, kind): the real reference when oracle/_ref travelled here, else the C oracle port."""
    from oracle import pyoracle
    Ni, Nj, Nk = shape
    if pyoracle.have_reference():
        # all the host threads: torchrun exports OMP_NUM_THREADS=1 to its children, which would silently make this a
        # single-thread baseline at N > 1
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        if pyoracle.Reference.max_threads() < ncpu:
            pyoracle.Reference.set_threads(ncpu)
        return pyoracle.Reference(Ni, Nj, Nk, C, C, C, 0.2, pml_percent=pml), "reference", pyoracle.Reference.max_threads()
    return pyoracle.Oracle(Ni, Nj, Nk, C, C, C, 0.2, pml_percent=pml), "port", (os.cpu_count() or 1)


def time_cpu(n, steps, warmup, pml=None, budget_s=25.0):
    """Gcell-updates/s of the CPU path on a bounded sample: the same 512x512 planes, fewer of them."""
    nk = n
    solver, kind, cores = cpu_solver((n, n, 32), pml)
    rng = np.random.default_rng(42)
    for c in range(6):
        solver.field(c)[...] = rng.uniform(-1, 1, size=(32, n, n))
    solver.update_fields()
    t0 = time.perf_counter()
    solver.update_fields()
    per_plane = (time.perf_counter() - t0) / 32.0
    solver.close()
    # planes such that (warmup + steps) steps take about `budget_s`
    nk = int(budget_s / max(per_plane * (steps + warmup), 1e-9))
    nk = max(32, min(n, (nk // 32) * 32))
    solver, kind, cores = cpu_solver((n, n, nk), pml)
    for c in range(6):
        a = solver.field(c)
        for k0 in range(0, nk, 32):
            a[k0:k0 + 32] = rng.uniform(-1, 1, size=(min(32, nk - k0), n, n))
    lo, hi, w, amp = sample_source_tables((n, n, nk), steps + warmup)
    jx, jy, jz = solver.field(6), solver.field(7), solver.field(8)

    def one(t):
        for kk in range(lo[2], hi[2]):
            for jj in range(lo[1], hi[1]):
                for ii in range(lo[0], hi[0]):
                    v = ((amp[t] * w[0][ii - lo[0]]) * w[1][jj - lo[1]]) * w[2][kk - lo[2]]
                    jx[kk, jj, ii] = v; jy[kk, jj, ii] = v; jz[kk, jj, ii] = v
        solver.update_fields()

    for t in range(warmup):
        one(t)
    t0 = time.perf_counter()
    for t in range(steps):
        one(warmup + t)
    dt = time.perf_counter() - t0
    solver.close()
    cells = n * n * nk
    return dict(value=cells * steps / dt / 1e9, unit="Gcell-updates/s", cores=cores, kind=kind,
                sample=f"{n}x{n}x{nk} fp64 periodic slab of the {n}^3 workload, {steps} steps after {warmup} warm-up, "
                       f"{'FDTD_openmp::FDTD (oracle/_ref)' if kind == 'reference' else 'C oracle port (oracle/liboracle.so)'}, "
                       f"OMP threads={cores}",
                ms_per_step=dt / steps * 1e3, cells=cells)


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = a.n
    r = time_cpu(n, a.steps, a.warmup, pml=(0.0625 if a.workload == "pml" else None), budget_s=100.0)
    line = {
        "impl": "reference", "metric": "Gcell-updates/s (E+B step)", "value": r["value"], "unit": "Gcell-updates/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(a, a.gpus, note="CPU arm runs a bounded slab of the same planes"),
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "Gcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(a, world, note=None):
    strong = getattr(a, "scaling", "weak") == "strong"
    cfg = {"workload": f"{a.n}^3 {a.dtype} Yee grid {'in total (z-slabs of n/N planes, BASELINE configs[3])' if strong else 'per GPU'}, {'periodic' if a.workload == 'periodic' else 'PML 32 cells (pml_percent 0.0625 in i/j, explicit thickness)'}"
                       f", random E/B seed 42 + sample.cpp point current source active every step (BASELINE configs[2])",
           "grid": [a.n, a.n, a.n * world if getattr(a, "scaling", "weak") == "weak" else a.n], "decomposition": f"z-slab x{world}" if world > 1 else "single GPU",
           "dx=dy=dz": "C", "dt": 0.2, "l2": "working set 12+ GiB per GPU >> 126 MB L2 (no flush needed)"}
    if note:
        cfg["note"] = note
    return cfg


# --------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist

    import fdtd_method_b200 as fb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- fdtd_method_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.n
    dtype = np.float64 if a.dtype == "f64" else np.float32
    W = 8 if a.dtype == "f64" else 4
    Nk = n * world if a.scaling == "weak" else n
    p = fb.Parameters(n, n, Nk, -n / 2 * C, n / 2 * C, -n / 2 * C, n / 2 * C, -Nk / 2 * C, Nk / 2 * C, C, C, C)
    kw = dict(dtype=dtype, device=local, rank=rank, nranks=world)
    if a.workload == "pml":
        g = fb.FDTD_PML(p, 0.2, pml_thickness=(32, 32, 32), **kw)
    else:
        g = fb.FDTD(p, 0.2, **kw)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(fb.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        g.comm_init(bytes(idt.cpu().numpy().tobytes()))
    info = g.info()
    nk_local = info.k_end - info.k_begin
    cells_local = n * n * nk_local
    cells_total = n * n * Nk

    # synthetic inputs in pinned host memory (also the e2e leg's H2D source)
    tdt = torch.float64 if a.dtype == "f64" else torch.float32
    host, pinned = [], True
    for _ in range(6):
        t = torch.empty((nk_local, n, n), dtype=tdt)
        if pinned:
            try:
                t = t.pin_memory()
            except Exception as e:      # 6 x field per rank (48 GiB at 8 x 512^3 fp64): fall back to pageable rather than lose the run
                pinned = False
                print(f"bench.py: rank {rank}: pin_memory failed ({e!r}); the e2e leg uses pageable host buffers", file=sys.stderr)
        host.append(t)
    rng = np.random.default_rng(42 + rank)
    for t in host:
        v = t.numpy()
        for k0 in range(0, nk_local, 64):
            v[k0:k0 + 64] = rng.uniform(-1, 1, size=v[k0:k0 + 64].shape)
    total_steps = a.warmup + a.steps
    lo, hi, w, amp = sample_source_tables((n, n, Nk), total_steps + 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg: inputs already in HBM when the timed region starts --------------------------
    for c in range(6):
        g.upload(c, host[c].numpy())
    g.set_source(lo, hi, w[0], w[1], w[2], amp)
    g.step(a.warmup)
    g.sync()
    barrier()
    launches0 = g.info().launches
    passes0 = g.info().passes_t2
    sampler = ClockSampler(local)
    sampler.start()
    g.timer_start()
    g.step(a.steps)
    ms = g.timer_stop()
    clocks = sampler.result()
    barrier()
    launches = g.info().launches - launches0
    g_passes_t2 = g.info().passes_t2 - passes0
    g.sync()

    # ---- end-to-end leg: HOST buffers in, HOST buffers out, through the public API -------------------------
    # upload the 6 fields from pinned memory, every step write that step's J from the host (the reference
    # sample's `get_field(JX)[idx] = v`, sample.cpp:66-81) and read back a 10x10 Ex probe (sample.cpp:125-134),
    # then download the 6 fields.
    g.clear_source()
    g.zeroed_currents()
    src_idx = np.array([i + j * n + k * n * n for k in range(lo[2], hi[2]) for j in range(lo[1], hi[1])
                        for i in range(lo[0], hi[0])], dtype=np.int64)
    kmid = info.k_begin + nk_local // 2
    probe_idx = np.array([i + j * n + kmid * n * n for j in range(n // 2 - 5, n // 2 + 5) for i in range(n // 2 - 5, n // 2 + 5)],
                         dtype=np.int64)
    wprod = np.array([(w[0][(q % n) - lo[0]], w[1][((q // n) % n) - lo[1]], w[2][(q // (n * n)) - lo[2]]) for q in src_idx])
    e2e_steps = a.steps
    barrier()
    t0 = time.perf_counter()
    for c in range(0 if a.no_e2e else 6):
        g.upload(c, host[c].numpy())
    probe_sum = 0.0
    for t in range(0 if a.no_e2e else e2e_steps):
        vals = (((amp[t] * wprod[:, 0]) * wprod[:, 1]) * wprod[:, 2]).astype(dtype)
        for c in (6, 7, 8):
            g.scatter(c, src_idx, vals)
        g.update_fields()
        probe_sum += float(g.gather(0, probe_idx).sum())
    for c in range(0 if a.no_e2e else 6):
        g.download(c, host[c].numpy())
    barrier()
    e2e_s = time.perf_counter() - t0
    field_bytes = 6 * cells_local * W
    h2d = field_bytes / e2e_steps + 3 * (src_idx.nbytes + src_idx.size * W) + probe_idx.nbytes
    d2h = field_bytes / e2e_steps + probe_idx.size * W

    # ---- reduce over ranks ------------------------------------------------------------------------------------
    if world > 1:
        tt = torch.tensor([ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(tt[0]), float(tt[1])
        ll = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(ll, op=dist.ReduceOp.SUM)
        launches = int(ll[0])
    value = cells_total * a.steps / (ms * 1e-3) / 1e9
    e2e_value = None if a.no_e2e else cells_total * e2e_steps / e2e_s / 1e9

    if rank == 0:
        peak, peak_src = measured_peak()
        if a.workload == "pml":
            npml = cells_local - max(n - 64, 0) ** 2 * (max(nk_local - 64, 0) if world == 1 else nk_local)
            alg_bytes = ((cells_local - npml) * WORDS_PER_CELL_STEP + npml * WORDS_PER_PML_CELL_STEP) * W
        else:
            alg_bytes = cells_local * WORDS_PER_CELL_STEP * W
        fused = bool(info.fused)
        pml = a.workload == "pml"
        t2_passes = g_passes_t2                       # two-step passes inside the timed region
        t2 = t2_passes * 2 >= a.steps - 1
        # the dominant launch (group): one T2 launch = 2 Yee steps; one fused launch = 1 step; two sweeps = 1 step;
        # PML solver: T2 launch on the main box's core + 4 rim sweeps = 2 steps, or 4 sweeps = 1 step
        steps_per_launch = 2 if t2 else 1
        # average duration: CUDA events over the timed region on the solver's stream (launch gaps and the 3 us
        # source kernels are < 1 % of it, profiles/launches_r01.csv)
        kernel_ms = ms / (a.steps / steps_per_launch)
        achieved = alg_bytes * steps_per_launch / (kernel_ms * 1e-3) / 1e9
        if pml:
            main_words = 6 if t2 else 18
            moved_bytes = ((cells_local - npml) * main_words + npml * WORDS_PER_PML_CELL_STEP) * W   # per step
            kernel = ("PML step group: fused_BE_T2_kernel on the main box's core + 4 rim sweeps (sweep_B/E_kernel<PML>) = TWO Yee steps"
                      if t2 else "PML step: sweep_B/E_kernel interior + shell launches (4 launches = one Yee step)")
        else:
            moved_words = 6 if t2 else (12 if fused else 18)
            moved_bytes = cells_local * moved_words * W
            kernel = ("fused_BE_T2_kernel (one launch = TWO Yee steps of this rank's slab)" if t2 else
                      "fused_BE_kernel (one launch = one Yee step of this rank's slab)" if fused
                      else "sweep_B_kernel + sweep_E_kernel (two launches = one Yee step)")
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(a.dtype, n) if not pml else None,
                "peak_source": peak_src,
                "kernel": kernel,
                "kernel_ms": kernel_ms, "steps_per_launch": steps_per_launch,
                "algorithmic_bytes_per_cell_step": alg_bytes / cells_local,
                "kernel_compulsory_bytes_per_cell_step": moved_bytes / cells_local,
                "kernel_compulsory_GBs": moved_bytes * steps_per_launch / (kernel_ms * 1e-3) / 1e9,
                "note": "achieved = SURVEY.md 8(d)'s algorithmic bytes (21 words per interior cell-step, 36 per PML cell-step) x cells x "
                        "steps per launch / launch time; the temporally blocked pass really moves 6 words per cell-step (12 per "
                        "launch), so frac > 1 is expected and kernel_compulsory_GBs / traffic give the physical DRAM view "
                        "(DESIGN.md, 'Roofline accounting')"}
        line = {
            "metric": "Gcell-updates/s (E+B step)", "value": value, "unit": "Gcell-updates/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": a.scaling, "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
            "config": workload_config(a, world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Gcell-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": f"upload 6 fields from {'pinned' if pinned else 'PAGEABLE (pin_memory failed)'} host + {e2e_steps} x (scatter J from host, update_fields, gather "
                            f"10x10 Ex probe to host) + download 6 fields, wall clock, max over ranks", "probe_checksum": probe_sum},
            "gpu_launches": launches,
            "roofline": roof,
        }
        if world == 1 and not a.no_cpu:
            try:
                r = time_cpu(n, 4, 1, pml=None, budget_s=15.0)
                line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:  # the checker is optional for the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "Gcell-updates/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}
        emit(line)
    g.close()
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def emit(line: dict) -> None:
    """The one JSON line goes to the process's original stdout; everything else printed to fd 1 by libraries
    (NCCL's version banner, torchrun notices) was re-routed to stderr in main()."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--size", dest="n", type=int, default=512)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--workload", default="periodic", choices=["periodic", "pml"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the driver's contract): n^3 per GPU; strong: n^3 in total, z-slabs of n/N planes (BASELINE configs[3])")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (scaling probes)")
    a = ap.parse_args()
    if a.warmup < 3:
        a.warmup = 3
    if a.impl == "reference":
        run_reference_arm(a)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one process per GPU (the driver launches torchrun itself)
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd, stdout=_JSON_OUT))   # the children print the JSON line to our real stdout
    run_ours(a)


if __name__ == "__main__":
    main()
