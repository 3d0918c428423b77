"""torchrun worker for tests/test_multi_gpu.py: z-slab solvers on WORLD_SIZE GPUs vs the single-domain oracle.
Exit code 0 = every rank bit-identical to its slab of the oracle's fields."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import fdtd_method_b200 as fb  # noqa: E402
from fdtd_method_b200.slab import create_distributed  # noqa: E402
from oracle.pyoracle import C, J_KOKKOS, Oracle  # noqa: E402
from tests.util import params, seeded_fields  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    cases = [
        # (shape, pml, fusion, dtype, steps)
        ((16, 12, 10), None, True, np.float64, 7),
        ((16, 12, 10), None, False, np.float64, 7),
        ((32, 20, 2 * world), None, True, np.float64, 5),      # slabs of 2 planes
        ((32, 20, world), None, True, np.float64, 4),          # slabs of 1 plane
        ((64, 40, 33), None, True, np.float32, 6),
        ((20, 16, 24), 0.2, False, np.float64, 8),
        ((33, 9, 11), None, True, np.float64, 5),              # odd Ni -> two-sweep kernels
        ((32, 16, 32 * world), None, True, np.float64, 5),     # 32 planes per rank: overlapped halo exchange path
        ((64, 48, 40 * world), 0.1, False, np.float64, 5),     # PML interior/shell split on slabs
        ((32, 16, 8 * world), None, True, np.float64, 6),      # smallest slab that takes the overlapped path (H = 2)
        ((64, 24, 12 * world), None, True, np.float32, 5),     # overlapped path, fp32
        ((64, 48, 24 * world), 0.1, True, np.float64, 6),      # PML two-step pass on slabs: T2 core + rim sweeps + mid-pass exchanges
        ((64, 40, 20 * world), 0.2, True, np.float64, 5),      # ... k shell takes a large part of the first / last slab
        ((64, 9, 16 * world), 0.07, True, np.float64, 5),      # ... no shell in j, thin shells elsewhere
        # the benchmarked flavours on slab ranks: tiles wide enough for the TMA ring (Ni >= 122, Nj >= 26)
        # reading ghost planes through the tensor maps (wrap_k = 0), both storage types, several k chunks per slab
        # (chunk order 1..nz-1, 0 with the halo wait inside the kernel), and the PML two-step pass on top
        ((128, 48, 12 * world), None, True, np.float64, 6),
        ((192, 40, 40 * world), None, True, np.float64, 5),
        ((128, 48, 64 * world), None, True, np.float64, 4),
        ((136, 36, 24 * world), None, True, np.float32, 6),
        ((128, 64, 40 * world), 0.1, True, np.float64, 6),
        # uneven slabs around the two-step threshold: 5,4,.. planes -> every rank pairs; 4,..,3 -> nobody does
        ((32, 16, 4 * world + 1), None, True, np.float64, 5),
        ((32, 16, 4 * world - 1), None, True, np.float64, 5),
    ]
    if os.environ.get("MGPU_CASES"):
        sel = [int(v) for v in os.environ["MGPU_CASES"].split(",")]
        cases = [cases[i] for i in sel]
    failures = 0
    transport = in_kernel = None
    for shape, pml, fusion, dtype, steps in cases:
        Ni, Nj, Nk = shape
        d = (C, 1.25 * C, 0.8 * C)
        p = params(Ni, Nj, Nk, *d)
        if pml is None:
            g = create_distributed(fb.FDTD, p, 0.2, dtype=dtype, fusion=fusion)
        else:
            g = create_distributed(fb.FDTD_PML, p, 0.2, pml, dtype=dtype, fusion=fusion)
        o = Oracle(Ni, Nj, Nk, d[0], d[1], d[2], 0.2, dtype=dtype, j_mode=J_KOKKOS, pml_percent=pml)
        f = seeded_fields(23, (Nk, Nj, Ni), dtype=dtype, same_j=False)
        kb, ke = g.k_begin, g.k_end
        transport, in_kernel = g.info().transport, g.info().halo_in_kernel
        for c in range(9):
            o.field(c)[...] = f[c]
            g.upload(c, f[c][kb:ke])
        for t in range(steps):
            o.update_fields()
            g.update_fields()
            if t == 0 and pml is None:
                # collective slice read of a B component after an odd step: every rank flushes the deferred
                # half step (ring exchange), the owner of the plane returns it
                kq = Nk // 2
                sl = g.read_slice(4, 2, kq)
                own = kb <= kq < ke
                if (sl is not None) != own or (own and not np.array_equal(sl, o.field(4)[kq])):
                    failures += 1
                    print(f"[rank {rank}] MISMATCH read_slice(By, k={kq}) {shape}", flush=True)
            if t == 1:   # mid-run read forces the deferred half step + ghost refresh path
                for c in range(6):
                    if not np.array_equal(g.download(c), o.field(c)[kb:ke]):
                        failures += 1
                        print(f"[rank {rank}] MISMATCH mid-run {shape} pml={pml} fusion={fusion} comp {c}", flush=True)
        # reference-style loop with a J write before every call and no read in between (sample.cpp:66-87): the writes that
        # arrive while a call is recorded ride into the second stage of the pair as a pending box -- every rank gets the
        # same global index list; the box straddles the boundary between rank 0 and rank 1
        if pml is None or fusion:
            kq = ke0 = (Nk // world) if world > 1 else Nk // 2
            jidx = np.array([i + j * Ni + k * Ni * Nj for k in (kq - 1, kq) for j in (2, 3) for i in (4, 5, 6)])
            jidx = jidx[(jidx >= 0) & (jidx < Ni * Nj * Nk)]
            for t in range(4):
                for c in (6, 7, 8):
                    v = (np.arange(jidx.size) * 0.01 + 0.1 * (t + 1) * (c - 5)).astype(dtype)
                    g.scatter(c, jidx, v)
                    o.field(c).reshape(-1)[jidx] = v
                o.update_fields()
                g.update_fields()
        # batched steps: fdtd_step(n) pairs steps into the temporally blocked T2 pass (two ghost planes per side,
        # J ghost planes included) wherever the slab has >= 4 planes
        o.step(steps + 1)
        g.step(steps + 1)
        for c in range(6):
            if not np.array_equal(g.download(c), o.field(c)[kb:ke]):
                failures += 1
                print(f"[rank {rank}] MISMATCH {shape} pml={pml} fusion={fusion} {np.dtype(dtype).name} comp {c}", flush=True)
        g.close()
    t = torch.tensor([failures], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print(f"mgpu_worker: world={world} cases={len(cases)} transport={transport} halo_in_kernel={in_kernel} total mismatches={int(t[0])}", flush=True)
    dist.destroy_process_group()
    sys.exit(1 if int(t[0]) else 0)


if __name__ == "__main__":
    main()
