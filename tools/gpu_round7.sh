#!/bin/bash
out=gpurun_out/${1:-r01m}; mkdir -p $out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1
tail -4 $out/pytest_gpu.log | cut -c1-300
timeout 900 python tools/config_table.py > $out/config_table.jsonl 2> $out/config_table.err
cat $out/config_table.jsonl
timeout 600 python bench.py --steps 200 --warmup 10 > $out/bench_n1.json 2> $out/bench_n1.err; cat $out/bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > $out/bench_under_ncu.log 2>&1
