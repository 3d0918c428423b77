"""Quick GPU check of the T2 pass against the one-step fused pass on arbitrary shapes (bit equality + timing).
    python tools/t2_check.py Ni Nj Nk steps [dtype]
Env: FDTD_B200_T2_VARIANT, FDTD_B200_FUSED_KC."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd_method_b200 as fb
C = 3e10
Ni, Nj, Nk, steps = map(int, sys.argv[1:5])
dtype = np.float32 if len(sys.argv) > 5 and sys.argv[5] == "f32" else np.float64
p = fb.Parameters(Ni, Nj, Nk, 0, Ni * C, 0, Nj * C, 0, Nk * C, C, C, C)
rng = np.random.default_rng(1)
f = [rng.uniform(-1, 1, size=(Nk, Nj, Ni)).astype(dtype) for _ in range(6)]
res = []
for temporal in (False, True):
    g = fb.FDTD(p, 0.2, dtype=dtype, temporal=temporal)
    for c in range(6):
        g.upload(c, f[c])
    g.sync(); g.timer_start(); g.step(steps); ms = g.timer_stop()
    res.append([g.download(c) for c in range(6)])
    print(f"temporal={temporal} {ms/steps:.4f} ms/step launches={g.info().launches}", flush=True)
    g.close()
ok = all(np.array_equal(a, b) for a, b in zip(*res))
print("T2 == T1:", ok, flush=True)
sys.exit(0 if ok else 1)
