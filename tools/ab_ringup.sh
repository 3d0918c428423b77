# A/B of a build-time switch on one box: default library vs the alternative build (FDTD_B200_LIB), alternating runs
out=gpurun_out/${TAG:-r02j}; mkdir -p $out
ALT=${ALT:-/root/repo/fdtd_method_b200/libfdtd_b200_noringup.so}
( time timeout 900 python -m pytest tests -m gpu -q -x ) > $out/pytest_gpu.log 2>&1; tail -3 $out/pytest_gpu.log | cut -c1-200
B="--steps 200 --warmup 10 --reps 3 --no-e2e --no-cpu --no-verify"
for i in 1 2; do
  timeout 200 python bench.py $B > $out/f64_default_$i.json 2>> $out/err.log
  FDTD_B200_LIB=$ALT timeout 200 python bench.py $B > $out/f64_alt_$i.json 2>> $out/err.log
  timeout 200 python bench.py $B --dtype f32 --f32-arith > $out/f32a_default_$i.json 2>> $out/err.log
  FDTD_B200_LIB=$ALT timeout 200 python bench.py $B --dtype f32 --f32-arith > $out/f32a_alt_$i.json 2>> $out/err.log
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$out/f*_*.json")):
    try:
        d=json.loads(open(f).read()); print(f.split('/')[-1], round(d['value'],2), [round(x,1) for x in d['rep_ms']], d['clocks'].get('sm_mhz'))
    except Exception as e: print(f,'ERR',e)
PY
