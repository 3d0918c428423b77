"""Small driver for compute-sanitizer (tools/gpu_round.sh stage `sanitize`): runs every kernel family once at sizes the
tools finish in seconds -- T2 pass (TMA ring tiles and cp.async ring tiles, J tiles, fp64 / fp32 / fp32-arithmetic),
one-step fused pass, two sweeps, PML rim sweeps and the PML two-step pass, scatter / gather / slice / source kernels --
and checks the results against the CPU oracle, so a run that passes is a run whose results were right as well."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fdtd_method_b200 as fb  # noqa: E402
from oracle.pyoracle import C  # noqa: E402
from tests.util import assert_bit_equal, load_both, make_pair, seeded_fields  # noqa: E402

cases = [
    # shape, dtype, pml, f32_arith, steps
    ((128, 48, 8), np.float64, None, False, 5),     # TMA tiles + wrap tiles, a pair + an odd step (one-step fused pass)
    ((136, 36, 6), np.float32, None, False, 4),     # fp32 storage, 68-cell boxes
    ((256, 40, 6), np.float32, None, True, 5),      # fp32 arithmetic, 128-cell boxes, odd step through the sweeps
    ((33, 7, 5), np.float64, None, False, 3),       # odd Ni: sweep kernels only
    ((64, 48, 24), np.float64, 0.1, False, 5),      # PML two-step pass: T2 core + rim sweeps
    ((20, 16, 12), np.float64, 0.2, False, 3),      # PML sweeps (no room for the pair path)
]
only = os.environ.get("SANITIZE_CASES")
if only:
    cases = [cases[int(i)] for i in only.split(",")]
for shape, dtype, pml, fa, steps in cases:
    Ni, Nj, Nk = shape
    o, g = make_pair(Ni, Nj, Nk, d=(C, 1.25 * C, 0.8 * C), dtype=dtype, pml=pml, f32_arith=fa)
    f = seeded_fields(3, (Nk, Nj, Ni), dtype=dtype, same_j=False)
    load_both(o, g, f, comps=range(6))
    idx = np.array([5 + 3 * Ni + 2 * Ni * Nj, 6 + 3 * Ni + 2 * Ni * Nj])
    for c in (6, 7, 8):
        g.scatter(c, idx, f[c].reshape(-1)[idx])
        o.field(c).reshape(-1)[idx] = f[c].reshape(-1)[idx]
    o.step(steps); g.step(steps)
    assert_bit_equal(o, g, what=f"sanitize {shape}")
    g.read_slice(4, 2, Nk // 2)
    g.gather(0, idx)
    o.update_fields(); g.update_fields()
    assert_bit_equal(o, g, what=f"sanitize {shape} +1")
    g.close()
    print("ok", shape, np.dtype(dtype).name, "pml" if pml else "periodic", "f32_arith" if fa else "", flush=True)
print("sanitize_driver: all cases bit-exact")
