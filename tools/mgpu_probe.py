"""Where the multi-GPU step loses time (run under torchrun on the GPU box, one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_probe.py

Times step(K) of the 512^3-per-GPU periodic workload (device events, max over ranks) under the timing-only
switches of FDTD_B200_MGPU_DEBUG: default (overlapped exchange), exchange on the compute stream, and the same two
launch structures with the exchange skipped (fields wrong, timing only) -- the differences are the cost of the NCCL
exchange, of the interior / boundary split, and of rank skew.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd_method_b200 as fb  # noqa: E402

C = 3e10


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(os.environ.get("PROBE_N", "512"))
    steps = int(os.environ.get("PROBE_STEPS", "60"))
    Nk = n * world
    p = fb.Parameters(n, n, Nk, 0, n * C, 0, n * C, 0, Nk * C, C, C, C)
    g = fb.FDTD(p, 0.2, device=local, rank=rank, nranks=world)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(fb.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        g.comm_init(bytes(idt.cpu().numpy().tobytes()))
    rng = np.random.default_rng(rank)
    plane = rng.uniform(-1, 1, size=g.local_shape)
    for c in range(6):
        g.upload(c, plane)
    rows = []
    variants = [("3 streams H=2 (default)", 0), ("3 streams H=4", 4 << 8), ("boundary after interior H=2", 4),
                ("boundary after interior H=4 (previous default)", 4 | (4 << 8)), ("no overlap", 1),
                ("3 streams, exchange skipped", 2), ("single launch, exchange skipped", 3)]
    if os.environ.get("PROBE_FIRST_ONLY"):
        variants = variants[:1]
    for name, dbg in variants:
        os.environ["FDTD_B200_MGPU_DEBUG"] = str(dbg)
        best = None
        for rep in range(3):
            g.step(6)
            g.sync()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            g.timer_start()
            g.step(steps)
            ms = g.timer_stop()
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0]) / steps
            best = ms if best is None else min(best, ms)
        rows.append({"variant": name, "ms_per_step": best, "gcells": n * n * Nk / best / 1e6, "n_gpus": world})
        if rank == 0:
            print(json.dumps(rows[-1]), flush=True)
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/mgpu_probe_n{world}.jsonl", "w") as fh:
            for r in rows:
                fh.write(json.dumps(r) + "\n")
    g.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
