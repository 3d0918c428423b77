#!/usr/bin/env python
"""bench.py -- Gcell-updates/s of the Yee time step (update_fields) on B200, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 512] [--dtype f64|f32]
                    [--workload periodic|pml]

One "step" = one update_fields() (B half, E, B half) over the whole grid.  Workload at N=1: the
configuration BASELINE.json's metric is quoted on, a 512^3 fp64 Yee grid, periodic boundaries, random
initial E/B (seed 42) plus the reference sample's point current source kept active (configs[2]); at N>1
the same 512^3 block PER GPU, z-slab partitioned (weak scaling, 512 x 512 x 512*N), one process per GPU,
halo planes over NCCL send/recv.  One JSON line on stdout (rank 0).

Timing: after W warm-up steps, R (--reps, default 5) blocks of K steps are timed one by one with CUDA events on the
solver's stream; every block ends with fdtd_flush(), so the closing B half step of its last update_fields() is INSIDE
the timed region.  `ms_per_step` / `value` are the median block (max over ranks per block); all blocks are listed.

`--impl reference` times the reference's own CPU implementation (oracle/_ref/libfdtd_ref.so = the unmodified
src/FDTD/FDTD.cpp compiled by oracle/Makefile; the C oracle port when that file is absent) on the host cores, each leg in
a child process started with OMP_NUM_THREADS / OMP_PROC_BIND=spread / OMP_PLACES=threads: plain
C++/OpenMP at all threads (the line's value), at 1 thread, and the Kokkos-OpenMP path (oracle/_ref/libfdtd_ref_kokkos.so).
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
    """(per-launch DRAM bytes of the dominant kernel, where the number comes from): read from the committed ncu
    --set full capture (profiles/ncu_traffic.json, which names the commit and capture it was taken from) -- DRAM
    counters cannot be read live outside a profiler."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            d = json.load(fh)
        e = d.get(f"{dtype}_{n}", {})
        return e.get("dram_bytes_per_launch"), f"profiles/ncu_traffic.json: {e.get('source', d.get('source', 'ncu --set full'))}"
    except Exception:
        return None, None


def verify_against_cpu(fb, n, dtype, device):
    """4 steps of an n x n x 64 slab of the bench workload (random E/B seed 42 + the sample source) on the GPU against the
    CPU checker, bit for bit.  The checker is test infrastructure (oracle/): it verifies, it is never what is timed."""
    from oracle import pyoracle
    nk, steps = 64, 4
    p = fb.Parameters(n, n, nk, -n / 2 * C, n / 2 * C, -n / 2 * C, n / 2 * C, -nk / 2 * C, nk / 2 * C, C, C, C)
    g = fb.FDTD(p, 0.2, dtype=dtype, device=device, j_openmp_quirk=True)
    if dtype == np.float64 and pyoracle.have_reference():
        chk, kind = pyoracle.Reference(n, n, nk, C, C, C, 0.2), "reference (oracle/_ref, FDTD_openmp::FDTD)"
    else:
        chk, kind = pyoracle.Oracle(n, n, nk, C, C, C, 0.2, dtype=dtype, j_mode=pyoracle.J_OPENMP), "oracle port (oracle/liboracle.so)"
    rng = np.random.default_rng(42)
    for c in range(6):
        f = rng.uniform(-1, 1, size=(nk, n, n)).astype(dtype)
        chk.field(c)[...] = f
        g.upload(c, f)
    lo, hi, w, amp = sample_source_tables((n, n, nk), steps)
    g.set_source(lo, hi, w[0], w[1], w[2], amp)
    g.step(steps)
    for t in range(steps):
        for kk in range(lo[2], hi[2]):
            for jj in range(lo[1], hi[1]):
                for ii in range(lo[0], hi[0]):
                    v = ((amp[t] * w[0][ii - lo[0]]) * w[1][jj - lo[1]]) * w[2][kk - lo[2]]
                    for c in (6, 7, 8):
                        chk.field(c)[kk, jj, ii] = v
        chk.update_fields()
    worst = 0.0
    for c in range(6):
        worst = max(worst, float(np.abs(g.download(c).astype(np.float64) - chk.field(c).astype(np.float64)).max()))
    passes = int(g.info().passes_t2)
    g.close(); chk.close()
    return {"grid": [n, n, nk], "steps": steps, "checker": kind, "max_abs_diff": worst, "bit_exact": worst == 0.0, "t2_passes": passes}


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
            time.sleep(0.004)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------------
# CPU side: the reference's own implementation on the host cores
# --------------------------------------------------------------------------------------------------------
def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def cpu_leg_main(a):
    """Child process of one CPU leg (bench.py --cpu-leg openmp|kokkos ...): started with the OpenMP environment already
    in place, so libgomp reads OMP_NUM_THREADS / OMP_PROC_BIND / OMP_PLACES when it loads.  Prints one JSON object."""
    from oracle import pyoracle
    n, nk = a.n, a.nk
    pml = 0.0625 if a.workload == "pml" else None
    if a.cpu_leg == "kokkos":
        solver, kind = pyoracle.ReferenceKokkos(n, n, nk, C, C, C, 0.2, pml_percent=pml), "reference"
        threads = pyoracle.ReferenceKokkos.threads()
        what = "FDTD_kokkos::FDTD, Kokkos OpenMP backend (oracle/_ref/libfdtd_ref_kokkos.so)"
    elif pyoracle.have_reference():
        solver, kind = pyoracle.Reference(n, n, nk, C, C, C, 0.2, pml_percent=pml), "reference"
        threads = pyoracle.Reference.max_threads()
        what = "FDTD_openmp::FDTD (oracle/_ref/libfdtd_ref.so)"
    else:
        solver, kind = pyoracle.Oracle(n, n, nk, C, C, C, 0.2, pml_percent=pml), "port"
        threads = int(os.environ.get("OMP_NUM_THREADS", "1"))
        what = "C oracle port (oracle/liboracle.so)"
    rng = np.random.default_rng(42)
    for c in range(6):
        f = solver.field(c)
        for k0 in range(0, nk, 32):
            f[k0:k0 + 32] = rng.uniform(-1, 1, size=f[k0:k0 + 32].shape)
    total = a.warmup + a.steps * a.reps
    lo, hi, w, amp = sample_source_tables((n, n, nk), total)
    jx, jy, jz = solver.field(6), solver.field(7), solver.field(8)

    def one(t):   # the reference caller's per-step J writes (sample.cpp:66-81), then update_fields()
        for kk in range(lo[2], hi[2]):
            for jj in range(lo[1], hi[1]):
                for ii in range(lo[0], hi[0]):
                    v = ((amp[t] * w[0][ii - lo[0]]) * w[1][jj - lo[1]]) * w[2][kk - lo[2]]
                    jx[kk, jj, ii] = v; jy[kk, jj, ii] = v; jz[kk, jj, ii] = v
        solver.update_fields()

    t = 0
    for _ in range(a.warmup):
        one(t); t += 1
    blocks = []
    for _ in range(a.reps):
        t0 = time.perf_counter()
        for _ in range(a.steps):
            one(t); t += 1
        blocks.append((time.perf_counter() - t0) / a.steps)
    solver.close()
    med = float(np.median(blocks))
    emit(dict(ok=True, kind=kind, what=what, threads=threads, n=n, nk=nk, steps=a.steps, reps=a.reps, warmup=a.warmup,
                          s_per_step=med, s_per_step_all=blocks, cells=n * n * nk, value=n * n * nk / med / 1e9))


def cpu_leg(leg, threads, n, nk, steps, warmup, reps, workload="periodic", timeout=300):
    """Run one CPU leg in a child process with the OpenMP environment of BASELINE.md section 3; returns its JSON dict."""
    import subprocess
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="spread", OMP_PLACES="threads")
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-leg", leg, "--n", str(n), "--nk", str(nk), "--steps", str(steps),
           "--warmup", str(warmup), "--reps", str(reps), "--workload", workload]
    try:
        r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode == 0 and lines:
            return json.loads(lines[-1])
        return dict(ok=False, error=(r.stderr or r.stdout)[-300:])
    except Exception as e:
        return dict(ok=False, error=repr(e))


def planes_for_budget(n, per_plane_step_s, steps_total, budget_s):
    nk = int(budget_s / max(per_plane_step_s * steps_total, 1e-9))
    return max(32, min(n, (nk // 32) * 32))


def cpu_baseline(n, steps, warmup, reps, budget_s, workload="periodic", with_c1=False):
    """The reference's CPU paths on this box's host cores, on a bounded slab (n x n x nk planes) of the n^3 workload:
    plain C++/OpenMP at all threads (the headline CPU number), the same at 1 thread, and Kokkos-OpenMP at all threads.
    Median of `reps` blocks of `steps` steps after `warmup` steps."""
    ncpu = host_threads()
    probe = cpu_leg("openmp", ncpu, n, 32, 1, 1, 1, workload)
    if not probe.get("ok"):
        raise RuntimeError(f"cpu leg failed: {probe.get('error')}")
    per_plane = probe["s_per_step"] / 32.0
    total = warmup + steps * reps
    nk = planes_for_budget(n, per_plane, total, budget_s * 0.5)
    main = cpu_leg("openmp", ncpu, n, nk, steps, warmup, reps, workload)
    if not main.get("ok"):
        raise RuntimeError(f"cpu leg failed: {main.get('error')}")
    # 1 thread: ~ncpu x slower per plane -> a thinner slab; Kokkos: the same slab as the OpenMP leg
    nk1 = planes_for_budget(n, per_plane * ncpu * 0.7, total, budget_s * 0.25)
    one = cpu_leg("openmp", 1, n, nk1, steps, warmup, reps, workload)
    from oracle import pyoracle
    kok = cpu_leg("kokkos", ncpu, n, planes_for_budget(n, per_plane * 1.4, total, budget_s * 0.25), steps, warmup, reps, workload) \
        if (pyoracle.have_reference_kokkos() and workload == "periodic") else dict(ok=False, error="oracle/_ref/libfdtd_ref_kokkos.so not present")
    out = dict(value=main["value"], unit="Gcell-updates/s", cores=main["threads"], kind=main["kind"],
               sample=f"{n}x{n}x{main['nk']} fp64 {workload} slab of the {n}^3 workload (same planes, fewer of them), median of {reps} x {steps} steps "
                      f"after {warmup} warm-up, {main['what']}, OMP_NUM_THREADS={main['threads']} OMP_PROC_BIND=spread OMP_PLACES=threads",
               ms_per_step=main["s_per_step"] * 1e3, cells=main["cells"], cpu_model=cpu_model(), nproc=ncpu,
               build="g++ -O3 -fopenmp -DNDEBUG (the reference's flags, no -march)",
               one_thread=(dict(value=one["value"], cores=1, sample=f"{n}x{n}x{one['nk']} slab, {one['what']}") if one.get("ok") else dict(value=None, error=one.get("error"))),
               kokkos_openmp=(dict(value=kok["value"], cores=kok["threads"], sample=f"{n}x{n}x{kok['nk']} slab, {kok['what']}") if kok.get("ok") else dict(value=None, error=kok.get("error"))))
    if with_c1:
        # C1, the reference's own perf test at the CI size (`sample 512 25`: zero fields, point source, 25 steps), slab-bounded
        c1 = cpu_leg("openmp", ncpu, n, planes_for_budget(n, per_plane, 25, 12.0), 25, 0, 1, workload)
        out["c1_sample_25_steps"] = dict(value=c1.get("value"), sample=f"{n}x{n}x{c1.get('nk')} slab, 25 steps, no warm-up (sample.cpp's timed loop)") if c1.get("ok") else dict(value=None, error=c1.get("error"))
    return out


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_baseline(a.n, a.steps, a.warmup, max(a.reps, 3) if a.reps_given else 3, budget_s=90.0, workload=a.workload, with_c1=True)
    line = {
        "impl": "reference", "metric": "Gcell-updates/s (E+B step)", "value": r["value"], "unit": "Gcell-updates/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(a, a.gpus, note="CPU arm: a bounded slab of the same n x n planes (throughput per cell does not depend on the plane count); "
                                                    "at N > 1 the GPU arm's grid grows to n x n x n*N while this arm stays on one host"),
        "cpu_baseline": r,
        "e2e": {"value": r["value"], "unit": "Gcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(a, world, note=None):
    strong = getattr(a, "scaling", "weak") == "strong"
    cfg = {"workload": f"{a.n}^3 {a.dtype} Yee grid {'in total (z-slabs of n/N planes, BASELINE configs[3])' if strong else 'per GPU'}, {'periodic' if a.workload == 'periodic' else 'PML 32 cells (pml_percent 0.0625 in i/j, explicit thickness)'}"
                       f", random E/B seed 42 + sample.cpp point current source active every step (BASELINE configs[2])",
           "grid": [a.n, a.n, a.n * world if getattr(a, "scaling", "weak") == "weak" else (getattr(a, "planes", 0) or a.n)], "decomposition": f"z-slab x{world}" if world > 1 else "single GPU",
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
    Nk = n * world if a.scaling == "weak" else (a.planes or n)
    p = fb.Parameters(n, n, Nk, -n / 2 * C, n / 2 * C, -n / 2 * C, n / 2 * C, -Nk / 2 * C, Nk / 2 * C, C, C, C)
    kw = dict(dtype=dtype, device=local, rank=rank, nranks=world, f32_arith=bool(a.f32_arith))
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
    info = info2 = g.info()
    nk_local = info.k_end - info.k_begin
    cells_local = n * n * nk_local
    cells_total = n * n * Nk

    # synthetic inputs in pinned host memory (also the e2e leg's H2D source)
    tdt = torch.float64 if a.dtype == "f64" else torch.float32
    host, pinned = [], True
    for _ in range(0 if a.zero_init else 6):
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
    total_steps = a.warmup + a.steps * a.reps
    lo, hi, w, amp = sample_source_tables((n, n, Nk), total_steps + 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity spot check at the bench shape (rank 0, N = 1): 4 steps of a 64-plane slab of this very workload against
    # the reference itself (oracle/_ref) or the C oracle, bit for bit, BEFORE anything is timed ---------------------------
    verify = None
    if a.verify and world == 1 and a.workload == "periodic" and not a.f32_arith:
        verify = verify_against_cpu(fb, n, dtype, local)

    # ---- device-resident leg: inputs already in HBM when the timed region starts --------------------------
    for c in range(0 if a.zero_init else 6):
        g.upload(c, host[c].numpy())
    g.set_source(lo, hi, w[0], w[1], w[2], amp)
    g.step(a.warmup)
    g.sync()
    if a.timeline and world > 1:
        g.timeline_enable(a.reps * (a.steps // 2 + 1))
    barrier()
    launches0 = g.info().launches
    passes0 = g.info().passes_t2
    sampler = ClockSampler(local)
    sampler.start()
    rep_ms, rep_ms_passes = [], []
    for _ in range(a.reps):
        # one block = K update_fields() = the passes + the closing B half step (flush); both are timed, separately, so that
        # the roofline can use the passes alone while `value` is charged the whole block
        g.timer_start()
        g.step(a.steps)
        ms_passes = g.timer_stop()       # start event -> end of the passes
        g.flush()
        rep_ms.append(g.timer_stop())    # same start event -> end of the closing half step (the host gap in between is included)
        rep_ms_passes.append(ms_passes)
        barrier()
    clocks = sampler.result()
    launches = (g.info().launches - launches0) // a.reps          # per block
    g_passes_t2 = (g.info().passes_t2 - passes0) // a.reps
    g.sync()
    timeline = g.timeline_read().tolist() if (a.timeline and world > 1) else None

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
    if a.zero_init:
        a.no_e2e = True
    e2e_steps = a.steps
    amp = amp[-(e2e_steps + 1):]
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
    rep_ms_rank = list(rep_ms)
    if world > 1:
        tt = torch.tensor(rep_ms + rep_ms_passes + [e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)          # per block: the slowest rank
        rep_ms, rep_ms_passes, e2e_s = [float(v) for v in tt[:a.reps]], [float(v) for v in tt[a.reps:2 * a.reps]], float(tt[-1])
        ll = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(ll, op=dist.ReduceOp.SUM)
        launches = int(ll[0])
        cm = torch.tensor([clocks.get("sm_mhz") or 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(cm, op=dist.ReduceOp.MIN)
        clocks["sm_mhz_min_over_ranks"] = float(cm[0])
        if timeline is not None:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", f"timeline_n{world}_rank{rank}.json"), "w") as fh:
                json.dump({"rank": rank, "world": world, "columns": ["pass_start_ms", "halo_copies_start_ms", "halo_copies_done_ms", "pass_end_ms"],
                           "rep_ms_this_rank": rep_ms_rank, "passes": timeline}, fh)
    ms = float(np.median(rep_ms))
    ms_passes = float(np.median(rep_ms_passes))
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
        kernel_ms = ms_passes / (a.steps / steps_per_launch)   # the passes alone (the block's closing half-step sweep is a different kernel)
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
        traffic, traffic_src = ncu_traffic("f32_arith" if a.f32_arith else a.dtype, n) if not pml else (None, None)
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
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
            "reps": a.reps, "rep_ms": rep_ms, "rep_ms_passes_only": rep_ms_passes, "timing": "median of `reps` blocks of `steps` steps, each block timed with CUDA events on the solver's stream "
                                                        "(max over ranks per block) and closed by fdtd_flush(): the trailing B half step is inside the timed region "
                                                        "(`rep_ms` = passes + closing half step; `rep_ms_passes_only` feeds roofline.kernel_ms)",
            "scaling": a.scaling, "vs_baseline": None, "dtype": a.dtype + ("_arith32 (opt-in FDTD_FLAG_F32_ARITH, not a reference mode)" if a.f32_arith else ""),
            "data": "synthetic (zero initial fields + the sample source; no host arrays)" if a.zero_init else "synthetic",
            "config": workload_config(a, world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Gcell-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": f"upload 6 fields from {'pinned' if pinned else 'PAGEABLE (pin_memory failed)'} host + {e2e_steps} x (scatter J from host, update_fields, gather "
                            f"10x10 Ex probe to host) + download 6 fields, wall clock, max over ranks", "probe_checksum": probe_sum},
            "gpu_launches": launches, "gpu_launches_all_reps": launches * a.reps,
            "roofline": roof,
        }
        if verify is not None:
            line["verify"] = verify
        if world > 1:
            line["halo"] = {"transport": {0: "none", 1: "nccl send/recv", 2: "copy engines into peer-mapped ghost planes"}[int(info2.transport)],
                            "wait_in_kernel": bool(info2.halo_in_kernel)}
        if world == 1 and not a.no_cpu:
            try:
                line["cpu_baseline"] = cpu_baseline(n, 2, 1, 3, budget_s=24.0)
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
    ap.add_argument("--f32-arith", action="store_true", help="with --dtype f32: the opt-in FDTD_FLAG_F32_ARITH mode (float storage AND float "
                                                             "arithmetic; not a reference mode -- labelled in `dtype` and `config`)")
    ap.add_argument("--planes", type=int, default=0, help="strong scaling: total number of k planes (default n); e.g. --size 1024 --planes 256 "
                                                          "--gpus 2 gives every rank the 128-plane slab of the 1024^3 / 8-GPU configuration")
    ap.add_argument("--zero-init", action="store_true", help="fields start at zero (plus the source): no host arrays, no uploads, no e2e leg "
                                                             "(1024^3 on one GPU: 48 GiB of host fields are not worth generating for a timing run, SURVEY.md 8(d) C4)")
    ap.add_argument("--reps", type=int, default=None, help="timed blocks of --steps steps (default 5); the median block is reported")
    ap.add_argument("--no-verify", dest="verify", action="store_false", help="skip the 4-step parity spot check against the CPU checker (N = 1)")
    ap.add_argument("--timeline", action="store_true", help="N > 1: record per-pass CUDA events on every rank -> gpurun_out/timeline_n<N>_rank<r>.json")
    ap.add_argument("--cpu-leg", default=None, choices=["openmp", "kokkos"], help=argparse.SUPPRESS)
    ap.add_argument("--nk", type=int, default=32, help=argparse.SUPPRESS)
    a = ap.parse_args()
    a.reps_given = a.reps is not None
    if a.reps is None:
        a.reps = 5
    a.reps = max(1, a.reps)
    if a.cpu_leg:
        cpu_leg_main(a)
        return
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
