"""z-slab decomposition: host-side logic of the multi-GPU mode (one process per GPU).

The reference's only distributed design is the coarray program coarray/fdtd.F90: 1-D decomposition along
z (:149-164), ring topology (:32-33), and per step two one-plane halo transfers -- Bx,By of the top owned
plane to the next image (:90-91) and Ex,Ey of the bottom owned plane to the previous image (:97-98).
Here the same ring runs between B200s over NVLink: by default every rank pushes its boundary planes into the neighbours'
ghost planes with the copy engines (csrc/peer_ring.cu; the T2 pass waits for them inside the kernel), with NCCL
send/recv (csrc/nccl_ring.cu) as bootstrap and fallback.  This module holds the pieces that do not need a GPU: the
plane ranges, the exchange plan, and the torch.distributed bootstrap that ships the NCCL unique id to every rank.
(One process driving several GPUs: fdtd_method_b200/multi.py.)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

from . import solver as _solver
from .structures import Component

EX, EY, EZ, BX, BY, BZ = (int(Component.EX), int(Component.EY), int(Component.EZ),
                          int(Component.BX), int(Component.BY), int(Component.BZ))


def slab_range(Nk: int, rank: int, nranks: int):
    """Planes [k_begin, k_end) owned by ``rank`` (remainder planes go to the low ranks)."""
    return _solver.slab_range(Nk, rank, nranks)


@dataclass(frozen=True)
class PlaneMove:
    component: int     # Component value
    src_plane: int     # local plane index on the SENDER (0 .. nk-1)
    dst_ghost: int     # ghost plane on the RECEIVER: -1 (below plane 0) or +1 (above plane nk-1)
    direction: int     # +1: sender -> rank+1, -1: sender -> rank-1 (ring, periodic in k)


def halo_plan(nk_local: int, fused: bool) -> List[PlaneMove]:
    """Planes one rank SENDS before a time step, in issue order (the receiver posts the mirror image).

    two-sweep path: before the B sweep  Ex,Ey bottom plane -> rank-1 (its top ghost; B update reads E[k+1]),
                    before the E sweep  Bx,By top plane    -> rank+1 (its bottom ghost; E update reads B[k-1]).
    fused pass:     one exchange per step -- Bx,By,Ex,Ey,Ez top plane -> rank+1 (bottom ghost: lets the
                    receiver rebuild B'(-1) itself) and Ex,Ey bottom plane -> rank-1 (top ghost)."""
    top, bottom = nk_local - 1, 0
    if fused:
        return ([PlaneMove(c, top, -1, +1) for c in (BX, BY, EX, EY, EZ)] +
                [PlaneMove(c, bottom, +1, -1) for c in (EX, EY)])
    return ([PlaneMove(c, bottom, +1, -1) for c in (EX, EY)] +
            [PlaneMove(c, top, -1, +1) for c in (BX, BY)])


JX, JY, JZ = int(Component.JX), int(Component.JY), int(Component.JZ)


def halo_plan_t2(nk_local: int) -> List[PlaneMove]:
    """Planes one rank SENDS before a two-step (T2) pass -- the plan csrc/fdtd_capi.cu::exchange_t2 executes.
    ``dst_ghost`` is the receiver's ghost plane: -2, -1 below plane 0; +1, +2 = planes nk, nk+1.

    to rank+1: E, B planes nk-1 and nk-2 -> its ghosts -1 and -2 (it rebuilds B1(-2), B1(-1), E1(-1) itself), and
               J plane nk-1 -> ghost -1 (E1(-1) contains the current term);
    to rank-1: E, B, J plane 0 -> its ghost +1 (it rebuilds B1(nk), E1(nk)), and Ex, Ey plane 1 -> ghost +2
               (B1(nk) reads E(nk+1))."""
    nk = nk_local
    up = [PlaneMove(c, nk - d, -d, +1) for c in (EX, EY, EZ, BX, BY, BZ) for d in (1, 2)]
    up += [PlaneMove(c, nk - 1, -1, +1) for c in (JX, JY, JZ)]
    down = [PlaneMove(c, 0, +1, -1) for c in (EX, EY, EZ, BX, BY, BZ, JX, JY, JZ)]
    down += [PlaneMove(c, 1, +2, -1) for c in (EX, EY)]
    return up + down


def create_distributed(cls, parameters, dt, *args, group=None, **kw):
    """Construct one slab solver per torch.distributed rank (``cls`` = FDTD or FDTD_PML) and initialise the
    NCCL ring: rank 0 creates the unique id, torch.distributed broadcasts it."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.cuda.current_device()
    g = cls(parameters, dt, *args, device=dev, rank=rank, nranks=world, **kw)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(_solver.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0, group=group)
        g.comm_init(bytes(idt.cpu().numpy().tobytes()))
    return g
