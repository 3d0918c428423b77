# A/B of a build-time switch on one box: default library vs an alternative build (ALT=<path to .so>, loaded through
# FDTD_B200_LIB), parity of the alternative first, then alternating timed runs.  Usage: TAG=... ALT=... bash tools/ab_build.sh
out=gpurun_out/${TAG:-ab}; mkdir -p $out
ALT=${ALT:?path of the alternative library}
( FDTD_B200_LIB=$ALT timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "temporal or tma or f32_arith or dtype or pml_two_step or baseline_size" ) > $out/pytest_alt.log 2>&1; tail -2 $out/pytest_alt.log | cut -c1-200
B="--steps 200 --warmup 10 --reps 3 --no-e2e --no-cpu --no-verify"
for i in 1 2; do
  timeout 200 python bench.py $B > $out/f64_default_$i.json 2>> $out/err.log
  FDTD_B200_LIB=$ALT timeout 200 python bench.py $B > $out/f64_alt_$i.json 2>> $out/err.log
  timeout 200 python bench.py $B --dtype f32 --f32-arith > $out/f32a_default_$i.json 2>> $out/err.log
  FDTD_B200_LIB=$ALT timeout 200 python bench.py $B --dtype f32 --f32-arith > $out/f32a_alt_$i.json 2>> $out/err.log
done
timeout 200 python bench.py $B --dtype f32 > $out/f32_default_1.json 2>> $out/err.log
FDTD_B200_LIB=$ALT timeout 200 python bench.py $B --dtype f32 > $out/f32_alt_1.json 2>> $out/err.log
python - <<PY
import json,glob
for f in sorted(glob.glob("$out/f*_*.json")):
    try:
        d=json.loads(open(f).read()); print(f.split('/')[-1], round(d['value'],2), [round(x,1) for x in d['rep_ms']], d['clocks'].get('sm_mhz'))
    except Exception as e: print(f,'ERR',e)
PY
