#!/bin/bash
# One gpurun call that refreshes the round's evidence: GPU parity suite, bench (both arms), BASELINE config table,
# ncu launch list and one `ncu --set full` capture of the dominant kernel.  Output under gpurun_out/<tag>/.
tag=${1:-r01}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi -L > $out/gpu.txt
nproc >> $out/gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1
tail -3 $out/pytest_gpu.log
timeout 600 python bench.py --steps 200 --warmup 10 > $out/bench_n1.json 2> $out/bench_n1.err
cat $out/bench_n1.json
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > $out/bench_ref.json 2> $out/bench_ref.err
timeout 900 python tools/config_table.py > $out/config_table.jsonl 2> $out/config_table.err
cat $out/config_table.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu > $out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_BE_T2 -s 4 -c 1 -f -o $out/t2_full \
    python bench.py --steps 12 --warmup 3 --no-cpu > $out/ncu_full.log 2>&1
ls -la $out
