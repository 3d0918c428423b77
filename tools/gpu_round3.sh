#!/bin/bash
# 1-GPU call: parity suite, config table, PML launch list + full capture of the shell sweeps, bench.
tag=${1:-r01d}
out=gpurun_out/$tag
mkdir -p $out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1
tail -3 $out/pytest_gpu.log
timeout 900 python tools/config_table.py --quick > $out/config_table.jsonl 2> $out/config_table.err
cat $out/config_table.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file $out/launches_pml.csv \
    python bench.py --workload pml --steps 4 --warmup 3 --no-cpu > $out/bench_pml_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_ -s 12 -c 4 -f -o $out/pml_full \
    python bench.py --workload pml --steps 4 --warmup 3 --no-cpu > $out/ncu_pml_full.log 2>&1
timeout 600 python bench.py --workload pml --steps 100 --warmup 10 --no-cpu > $out/bench_pml.json 2> $out/bench_pml.err
cat $out/bench_pml.json
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu > $out/bench_n1.json 2> $out/bench_n1.err
cat $out/bench_n1.json
ls -la $out
