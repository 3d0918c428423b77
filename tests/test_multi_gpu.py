"""Multi-GPU z-slab parity (needs >= 2 GPUs on the box; skipped otherwise): every rank's slab must equal the
single-domain oracle bit for bit, for the fused pass, the two-sweep kernels, fp32 and PML -- on every halo transport:
copy engines into peer-mapped ghost planes with the halo wait inside the T2 kernel (default), the same transport with the
three-stream boundary launches, and NCCL send/recv.  Plus the one-process / several-GPU ring (fdtd_comm_init_local)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MODES = {
    "peer-inkernel": {},
    "peer-3stream": {"FDTD_B200_HALO_IN_KERNEL": "0"},
    "nccl": {"FDTD_B200_TRANSPORT": "nccl"},
}
EXPECT = {"peer-inkernel": "transport=2 halo_in_kernel=1", "peer-3stream": "transport=2 halo_in_kernel=0", "nccl": "transport=1 halo_in_kernel=0"}


@pytest.mark.gpu
@pytest.mark.parametrize("world,mode", [(2, "peer-inkernel"), (2, "peer-3stream"), (2, "nccl"), (4, "peer-inkernel"), (4, "nccl"),
                                        (8, "peer-inkernel")])
def test_zslab_ring_bit_exact(world, mode, gpu_count):
    if gpu_count < world:
        pytest.skip(f"needs {world} GPUs, have {gpu_count}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29650 + world + 10 * list(MODES).index(mode)), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900,
                       env=dict(os.environ, **MODES[mode]))
    assert r.returncode == 0, r.stdout[-4000:]
    assert "total mismatches=0" in r.stdout
    assert EXPECT[mode] in r.stdout, r.stdout[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("ngpu", [2, 4])
def test_single_process_ring(ngpu, gpu_count):
    """FDTDMulti / FDTD_PML_Multi: one host thread, one solver per GPU, linked by fdtd_comm_init_local -- the reference's
    class interface (no ranks) on several GPUs, bit-identical to the single-domain oracle."""
    if gpu_count < ngpu:
        pytest.skip(f"needs {ngpu} GPUs, have {gpu_count}")
    import fdtd_method_b200 as fb
    from oracle.pyoracle import C, J_KOKKOS, Oracle
    from tests.util import params, seeded_fields
    for shape, pml, dtype in [((128, 48, 24 * ngpu), None, np.float64), ((64, 40, 10 * ngpu), None, np.float32),
                              ((64, 48, 24 * ngpu), 0.1, np.float64), ((32, 16, 2 * ngpu), None, np.float64)]:
        Ni, Nj, Nk = shape
        d = (C, 1.25 * C, 0.8 * C)
        p = params(Ni, Nj, Nk, *d)
        devs = list(range(ngpu))
        g = fb.FDTDMulti(p, 0.2, devices=devs, dtype=dtype) if pml is None else fb.FDTD_PML_Multi(p, 0.2, pml, devices=devs, dtype=dtype)
        assert all(i.transport == 2 for i in g.info())
        o = Oracle(Ni, Nj, Nk, d[0], d[1], d[2], 0.2, dtype=dtype, j_mode=J_KOKKOS, pml_percent=pml)
        f = seeded_fields(61, (Nk, Nj, Ni), dtype=dtype, same_j=False)
        for c in range(9):
            o.field(c)[...] = f[c]
            g.upload(c, f[c])
        for t in range(5):           # reference-style loop with a J write and a probe read per step
            idx = np.array([3 + 5 * Ni + (Nk // 2) * Ni * Nj, 7 + 2 * Ni + (Nk - 1) * Ni * Nj])
            v = np.array([0.25 * (t + 1), -0.5], dtype=dtype)
            g.get_field(6)[idx] = v
            o.field(6).reshape(-1)[idx] = v
            g.update_fields(); o.update_fields()
            probe = np.array([1 + Ni + k * Ni * Nj for k in range(Nk)])
            assert np.array_equal(g.get_field(0)[probe], o.field(0).reshape(-1)[probe]), f"{shape} probe step {t}"
        g.step(6); o.step(6)
        for c in range(6):
            assert np.array_equal(g.download(c), o.field(c)), f"{shape} pml={pml} comp {c}"
        g.close()


@pytest.mark.gpu
@pytest.mark.timeout(240, method="thread")   # a host-side deadlock on the ring blocks inside a CUDA call: fail, do not hang
@pytest.mark.parametrize("ngpu", [2])
def test_random_call_sequences_on_the_one_process_ring(ngpu, gpu_count):
    """The API fuzz of tests/test_parity_gpu.py on FDTDMulti / FDTD_PML_Multi: the same random call sequences with the grid
    spread over several GPUs in one process (every call fans out; J writes, reads and sources take global indices)."""
    if gpu_count < ngpu:
        pytest.skip(f"needs {ngpu} GPUs, have {gpu_count}")
    import fdtd_method_b200 as fb
    from oracle.pyoracle import J_KOKKOS, Oracle
    from tests.util import fuzz_call_sequence, params

    def make(Ni, Nj, Nk, d, dtype, pml, f32_arith):
        o = Oracle(Ni, Nj, Nk, d[0], d[1], d[2], 0.2, dtype=dtype, j_mode=J_KOKKOS, pml_percent=pml, f32_arith=f32_arith)
        p = params(Ni, Nj, Nk, *d)
        kw = dict(devices=list(range(ngpu)), dtype=dtype, f32_arith=f32_arith)
        g = fb.FDTDMulti(p, 0.2, **kw) if pml is None else fb.FDTD_PML_Multi(p, 0.2, pml, **kw)
        return o, g

    for seed in range(12):
        fuzz_call_sequence(seed, make=make, min_nk=4 * ngpu)
