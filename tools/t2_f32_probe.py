"""fp32 T2 pass on grids with wrap-free (TMA) tiles, against the oracle (debug helper).
    PROBE_SHAPE="Ni,Nj,Nk" PROBE_STEPS=4 python tools/t2_f32_probe.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd_method_b200 as fb
from tests.util import make_pair, load_both, seeded_fields, assert_bit_equal
Ni, Nj, Nk = (int(v) for v in os.environ.get("PROBE_SHAPE", "128,48,12").split(","))
steps = int(os.environ.get("PROBE_STEPS", "4"))
dtype = np.float32 if os.environ.get("PROBE_DTYPE", "f32") == "f32" else np.float64
o, g = make_pair(Ni, Nj, Nk, dtype=dtype)
load_both(o, g, seeded_fields(1, (Nk, Nj, Ni), dtype=dtype, same_j=False), comps=range(6))
g.step(steps); g.sync()
print("gpu done", g.info().passes_t2, flush=True)
o.step(steps)
assert_bit_equal(o, g, what=f"{dtype} T2 {Ni}x{Nj}x{Nk}")
print("ok")
