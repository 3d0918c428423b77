#!/bin/bash
out=gpurun_out/${1:-r01e}; mkdir -p $out
for v in 0 1 0 1; do
  FDTD_B200_ST_CS=$v timeout 300 python tools/sweep.py --no-sweeps --t2 0 --steps 60 2>&1 | sed "s/^{/{\"st_cs\": $v, /" | tee -a $out/stcs.jsonl
done
for v in 0 1; do
  FDTD_B200_ST_CS=$v timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_op_read_hit_rate.pct --clock-control none -k regex:fused_BE_T2 -s 3 -c 2 --csv --log-file $out/stcs_ncu_$v.csv python tools/sweep.py --no-sweeps --t2 0 --steps 6 > /dev/null 2>&1
  grep -v "^==" $out/stcs_ncu_$v.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tail -8
done
