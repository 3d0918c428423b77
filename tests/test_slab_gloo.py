"""world_size-2/3 gloo test of the z-slab halo protocol (host logic, no GPU).

Each rank owns a slab with one ghost plane on either side, moves exactly the planes listed by
fdtd_method_b200.slab.halo_plan() through torch.distributed (gloo) and advances its slab with a numpy
restatement of the two-sweep update; the gathered result must equal the single-domain oracle bit for bit.
This pins WHICH planes travel WHEN -- the same plan csrc/fdtd_capi.cu executes over NCCL on the GPUs."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sweep_B(f, c, n_half):
    """B half step(s) on planes 1..nk of ghosted arrays [nk+2, Nj, Ni] (FDTD.cpp:121-126)."""
    Ex, Ey, Ez, Bx, By, Bz = f
    cx, cy, cz = c
    s = slice(1, -1)
    ex, ey, ez = Ex[s], Ey[s], Ez[s]
    exk, eyk = Ex[2:], Ey[2:]
    hx = cz * (eyk - ey) - cy * (np.roll(ez, -1, axis=1) - ez)
    hy = cx * (np.roll(ez, -1, axis=2) - ez) - cz * (exk - ex)
    hz = cy * (np.roll(ex, -1, axis=1) - ex) - cx * (np.roll(ey, -1, axis=2) - ey)
    for _ in range(n_half):
        Bx[s] = Bx[s] + hx
        By[s] = By[s] + hy
        Bz[s] = Bz[s] + hz


def _sweep_E(f, c):
    """E full step, J = 0 (FDTD.cpp:85-93)."""
    Ex, Ey, Ez, Bx, By, Bz = f
    cx, cy, cz = c
    s = slice(1, -1)
    bx, by, bz = Bx[s], By[s], Bz[s]
    bxk, byk = Bx[:-2], By[:-2]
    Ex[s] = Ex[s] + (cy * (bz - np.roll(bz, 1, axis=1)) - cz * (by - byk))
    Ey[s] = Ey[s] + (cz * (bx - bxk) - cx * (bz - np.roll(bz, 1, axis=2)))
    Ez[s] = Ez[s] + (cx * (by - np.roll(by, 1, axis=2)) - cy * (bx - np.roll(bx, 1, axis=1)))


def _exchange(fields, moves, rank, world):
    """Execute a list of PlaneMove: send my plane, receive the mirror plane into my ghost."""
    nk = fields[0].shape[0] - 2
    for m in moves:
        dst = (rank + m.direction) % world
        src = (rank - m.direction) % world
        send = torch.from_numpy(np.ascontiguousarray(fields[m.component][1 + m.src_plane]))
        recv = torch.empty_like(send)
        ops = [dist.P2POp(dist.isend, send, dst), dist.P2POp(dist.irecv, recv, src)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        ghost = 0 if m.dst_ghost < 0 else nk + 1
        fields[m.component][ghost] = recv.numpy()


def _worker(rank, world, port, shape, steps, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fdtd_method_b200.slab import halo_plan, slab_range
    from tests.util import seeded_fields
    Ni, Nj, Nk = shape
    C = 3e10
    d, dt = (C, 1.25 * C, 0.8 * C), 0.2
    cE = tuple((C * dt) / v for v in d)
    cB = tuple((C * dt) / (2.0 * v) for v in d)
    kb, ke = slab_range(Nk, rank, world)
    nk = ke - kb
    full = seeded_fields(17, (Nk, Nj, Ni))
    f = [np.zeros((nk + 2, Nj, Ni)) for _ in range(6)]
    for c in range(6):
        f[c][1:-1] = full[c][kb:ke]
    plan = halo_plan(nk, fused=False)
    e_moves, b_moves = [m for m in plan if m.component < 3], [m for m in plan if m.component >= 3]
    pending = False
    for _ in range(steps):                       # the deferred-half-step state machine of fdtd_capi.cu
        _exchange(f, e_moves, rank, world)
        _sweep_B(f, cB, 2 if pending else 1)
        _exchange(f, b_moves, rank, world)
        _sweep_E(f, cE)
        pending = True
    _exchange(f, e_moves, rank, world)           # flush
    _sweep_B(f, cB, 1)
    q.put((rank, kb, ke, [a[1:-1].copy() for a in f]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,shape", [(2, (8, 6, 10)), (3, (6, 5, 7)), (2, (4, 4, 2))])
def test_slab_ring_matches_single_domain_oracle(world, shape):
    from oracle.pyoracle import Oracle
    from tests.util import seeded_fields
    Ni, Nj, Nk = shape
    C = 3e10
    steps = 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    parts = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = Oracle(Ni, Nj, Nk, C, 1.25 * C, 0.8 * C, 0.2)
    full = seeded_fields(17, (Nk, Nj, Ni))
    for c in range(6):
        o.field(c)[...] = full[c]
    o.step(steps)
    for rank, kb, ke, f in parts:
        for c in range(6):
            assert np.array_equal(f[c], o.field(c)[kb:ke]), f"rank {rank} component {c}"


def test_halo_plan_contents():
    from fdtd_method_b200.slab import halo_plan
    two = halo_plan(16, fused=False)
    assert [(m.component, m.src_plane, m.dst_ghost, m.direction) for m in two] == [
        (0, 0, +1, -1), (1, 0, +1, -1), (3, 15, -1, +1), (4, 15, -1, +1)]
    fused = halo_plan(16, fused=True)
    assert len(fused) == 7 and sum(m.direction == +1 for m in fused) == 5
