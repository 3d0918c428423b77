#!/bin/bash
# T2 tile rasterisation: DRAM traffic / duration (ncu, 1 launch) and free-running time vs strip width and chunk length
out=gpurun_out/${1:-r01l}; mkdir -p $out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -2
for cfg in "512 0 1" "512 0 2" "512 0 3" "512 0 99" "512 64 99" "1024 0 1" "1024 0 2" "1024 128 1" "1024 64 99" "1024 0 99" "256 0 1" "256 0 99"; do
  set -- $cfg
  export FDTD_B200_T2_STRIP=$3
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:fused_BE_T2 -s 3 -c 1 --csv --log-file $out/s_$1_$2_$3.csv python tools/sweep.py --n $1 --no-sweeps --t2 0 --kc $2 --steps 6 > /dev/null 2>&1
  free=$(timeout 300 python tools/sweep.py --n $1 --no-sweeps --t2 0 --kc $2 --steps 40 2>/dev/null | python -c "import sys,json; print(json.loads(sys.stdin.readline())['gcells'])")
  python - $out/s_$1_$2_$3.csv $1 $2 $3 $free <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5 and r[0].isdigit()]
n=int(sys.argv[2]); kc=int(sys.argv[3]); w=int(sys.argv[4]); free=float(sys.argv[5])
d={}
for r in rows: d.setdefault(r[0],{})[r[-3]]=float(r[-1].replace(',',''))
for k,v in d.items():
    rd,wr,t=v['dram__bytes_read.sum'],v['dram__bytes_write.sum'],v['gpu__time_duration.sum']
    print(f'{{"n": {n}, "kc": {kc}, "strip_w": {w}, "ncu_ms_per_launch": {t/1e6:.3f}, "read_GB": {rd/1e9:.2f}, "write_GB": {wr/1e9:.2f}, "x_compulsory": {(rd+wr)/(96*n**3):.3f}, "ncu_gcells": {2*n**3/t:.1f}, "free_running_gcells_40_steps": {free:.1f}}}')
PY
done | tee $out/strip.jsonl
