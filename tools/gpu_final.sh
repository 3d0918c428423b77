#!/bin/bash
# Final evidence of the round on one GPU: parity suite, smoke, both bench arms, launch list, full ncu of the T2 pass (fp64 and fp32).
out=gpurun_out/${1:-r01z}; mkdir -p $out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1; tail -3 $out/pytest_gpu.log
timeout 200 python __graft_entry__.py --smoke > $out/smoke.log 2>&1; tail -2 $out/smoke.log
timeout 400 python bench.py --steps 200 --warmup 10 > $out/bench_n1.json 2> $out/bench_n1.err; cut -c1-200 $out/bench_n1.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $out/bench_ref.json 2> $out/bench_ref.err; cut -c1-200 $out/bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > $out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_BE_T2 -s 4 -c 1 -f -o $out/t2_f64_full python bench.py --steps 12 --warmup 3 --no-cpu --no-e2e > $out/ncu_f64.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_BE_T2 -s 4 -c 1 -f -o $out/t2_f32_full python bench.py --dtype f32 --steps 12 --warmup 3 --no-cpu --no-e2e > $out/ncu_f32.log 2>&1
ls -la $out | head -20
