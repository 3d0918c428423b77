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


# ---- the two-step (T2) pass: two ghost planes per side, J planes included ------------------------------------------
def _t2_pass_numpy(f, cE, cB, cJ, n_half, G=2):
    """Two Yee steps on a slab with G = 2 ghost planes per side ([nk + 4, Nj, Ni] arrays, plane p at index p + 2), on the
    shrinking plane ranges of csrc/fused_kernel_t2.cuh: B1 on [-2, nk], E1 on [-1, nk], B2 on [-1, nk-1], E2 on [0, nk-1].
    Same J for both steps (static J)."""
    Ex, Ey, Ez, Bx, By, Bz, Jx, Jy, Jz = f
    nk = Ex.shape[0] - 2 * G

    def upd_B(lo, hi, halves):
        s, sk = slice(lo + G, hi + G + 1), slice(lo + G + 1, hi + G + 2)
        ex, ey, ez = Ex[s], Ey[s], Ez[s]
        hx = cB[2] * (Ey[sk] - ey) - cB[1] * (np.roll(ez, -1, axis=1) - ez)
        hy = cB[0] * (np.roll(ez, -1, axis=2) - ez) - cB[2] * (Ex[sk] - ex)
        hz = cB[1] * (np.roll(ex, -1, axis=1) - ex) - cB[0] * (np.roll(ey, -1, axis=2) - ey)
        for _ in range(halves):
            Bx[s] = Bx[s] + hx
            By[s] = By[s] + hy
            Bz[s] = Bz[s] + hz

    def upd_E(lo, hi):
        s, sk = slice(lo + G, hi + G + 1), slice(lo + G - 1, hi + G)
        bx, by, bz = Bx[s], By[s], Bz[s]
        Ex[s] = Ex[s] + (((cJ * Jx[s]) + cE[1] * (bz - np.roll(bz, 1, axis=1))) - cE[2] * (by - By[sk]))
        Ey[s] = Ey[s] + (((cJ * Jy[s]) + cE[2] * (bx - Bx[sk])) - cE[0] * (bz - np.roll(bz, 1, axis=2)))
        Ez[s] = Ez[s] + (((cJ * Jz[s]) + cE[0] * (by - np.roll(by, 1, axis=2))) - cE[1] * (bx - np.roll(bx, 1, axis=1)))

    upd_B(-2, nk, n_half)
    upd_E(-1, nk)
    upd_B(-1, nk - 1, 2)
    upd_E(0, nk - 1)


def _exchange_t2(fields, moves, rank, world, G=2):
    nk = fields[0].shape[0] - 2 * G
    ops, recvs = [], []
    for m in moves:
        dst, src = (rank + m.direction) % world, (rank - m.direction) % world
        send = torch.from_numpy(np.ascontiguousarray(fields[m.component][G + m.src_plane]))
        recv = torch.empty_like(send)
        ops += [dist.P2POp(dist.isend, send, dst), dist.P2POp(dist.irecv, recv, src)]
        recvs.append((m, recv))
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    for m, recv in recvs:
        ghost = G + m.dst_ghost if m.dst_ghost < 0 else G + nk - 1 + m.dst_ghost
        fields[m.component][ghost] = recv.numpy()


def _worker_t2(rank, world, port, shape, pairs, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fdtd_method_b200.slab import halo_plan_t2, slab_range
    from tests.util import seeded_fields
    Ni, Nj, Nk = shape
    C, PI = 3e10, 3.14159265358
    d, dt = (C, 1.25 * C, 0.8 * C), 0.2
    cE = tuple((C * dt) / v for v in d)
    cB = tuple((C * dt) / (2.0 * v) for v in d)
    cJ = -4.0 * PI * dt
    kb, ke = slab_range(Nk, rank, world)
    nk = ke - kb
    full = seeded_fields(29, (Nk, Nj, Ni), same_j=False)
    f = [np.zeros((nk + 4, Nj, Ni)) for _ in range(9)]
    for c in range(9):
        f[c][2:-2] = full[c][kb:ke]
    plan = halo_plan_t2(nk)
    for p in range(pairs):
        _exchange_t2(f, plan, rank, world)
        _t2_pass_numpy(f, cE, cB, cJ, 1 if p == 0 else 2)
    q.put((rank, kb, ke, [a[2:-2].copy() for a in f[:3]]))     # E is final after a pass (B carries the deferred half step)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,shape", [(2, (8, 6, 10)), (3, (6, 5, 13)), (2, (4, 4, 8))])
def test_t2_halo_plan_matches_single_domain_oracle(world, shape):
    """The planes halo_plan_t2 lists (two ghost planes per side, J included) are exactly what a two-step pass needs."""
    from oracle.pyoracle import J_KOKKOS, Oracle
    from tests.util import seeded_fields
    Ni, Nj, Nk = shape
    C = 3e10
    pairs = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_t2, args=(r, world, port, shape, pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    parts = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = Oracle(Ni, Nj, Nk, C, 1.25 * C, 0.8 * C, 0.2, j_mode=J_KOKKOS)
    full = seeded_fields(29, (Nk, Nj, Ni), same_j=False)
    for c in range(9):
        o.field(c)[...] = full[c]
    o.step(2 * pairs)
    for rank, kb, ke, f in parts:
        for c in range(3):
            assert np.array_equal(f[c], o.field(c)[kb:ke]), f"rank {rank} component {c}"


def test_halo_plan_t2_contents():
    from fdtd_method_b200.slab import halo_plan_t2
    plan = halo_plan_t2(16)
    assert len(plan) == 26                                          # the 26 planes of csrc/fdtd_capi.cu::exchange_t2
    assert sum(m.direction == +1 for m in plan) == 15 and sum(m.direction == -1 for m in plan) == 11
    assert {(m.component, m.src_plane, m.dst_ghost) for m in plan if m.component >= 6} == {
        (6, 15, -1), (7, 15, -1), (8, 15, -1), (6, 0, +1), (7, 0, +1), (8, 0, +1)}


def test_halo_plan_contents():
    from fdtd_method_b200.slab import halo_plan
    two = halo_plan(16, fused=False)
    assert [(m.component, m.src_plane, m.dst_ghost, m.direction) for m in two] == [
        (0, 0, +1, -1), (1, 0, +1, -1), (3, 15, -1, +1), (4, 15, -1, +1)]
    fused = halo_plan(16, fused=True)
    assert len(fused) == 7 and sum(m.direction == +1 for m in fused) == 5
