#!/bin/bash
# DRAM traffic and duration of the T2 pass vs chunk length (ncu, 2 launches each)
out=gpurun_out/${1:-r01j}; mkdir -p $out
for cfg in "512 32" "512 64" "512 171" "1024 32" "1024 64" "1024 128" "1024 512"; do
  set -- $cfg
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:fused_BE_T2 -s 2 -c 2 --csv --log-file $out/kc_$1_$2.csv python tools/sweep.py --n $1 --no-sweeps --t2 0 --kc $2 --steps 6 > /dev/null 2>&1
  python - $out/kc_$1_$2.csv $1 $2 <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5 and r[0].isdigit()]
n=int(sys.argv[2]); kc=int(sys.argv[3])
d={}
for r in rows: d.setdefault(r[0],{})[r[-3]]=float(r[-1].replace(',',''))
for k,v in d.items():
    rd,wr,t=v['dram__bytes_read.sum'],v['dram__bytes_write.sum'],v['gpu__time_duration.sum']
    print(f'{{"n": {n}, "kc": {kc}, "ms_per_launch": {t/1e6:.3f}, "read_GB": {rd/1e9:.2f}, "write_GB": {wr/1e9:.2f}, "x_compulsory": {(rd+wr)/(96*n**3):.3f}, "gcells": {2*n**3/t:.1f}}}')
PY
done | tee $out/kc_traffic.jsonl
