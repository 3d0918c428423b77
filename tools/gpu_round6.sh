#!/bin/bash
# 1-GPU call: full parity suite, smoke(), lazy pairing timing, one-step pass variants, C++ sample, 1024^3 traffic.
out=gpurun_out/${1:-r01i}; mkdir -p $out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1
tail -4 $out/pytest_gpu.log | cut -c1-300
timeout 300 python __graft_entry__.py --smoke > $out/smoke.log 2>&1; tail -3 $out/smoke.log
timeout 300 python tools/e2e_breakdown.py --steps 60 > $out/e2e_breakdown.json 2> $out/e2e_breakdown.err; cat $out/e2e_breakdown.json
timeout 600 python tools/sweep.py --no-sweeps --variants 5 28 26 24 --steps 40 > $out/t1_variants.jsonl 2>&1; sed -i 's/^/T1 /' $out/t1_variants.jsonl
FDTD_B200_NO_T2=1 timeout 600 python tools/sweep.py --variants 5 28 26 24 --steps 40 2>&1 | grep fused | tee $out/t1_variants.jsonl
make -C cpp > $out/make_cpp.log 2>&1
( cd cpp/bin && time ./sample_b200 256 200 ) > $out/sample_256_200.log 2>&1; tail -5 $out/sample_256_200.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:fused_BE_T2 -s 2 -c 2 --csv --log-file $out/t2_1024_ncu.csv python tools/sweep.py --n 1024 --no-sweeps --t2 0 --steps 6 > /dev/null 2>&1
grep -v "^==" $out/t2_1024_ncu.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tail -6
timeout 600 python bench.py --steps 200 --warmup 10 > $out/bench_n1.json 2> $out/bench_n1.err; cat $out/bench_n1.json | cut -c1-400
