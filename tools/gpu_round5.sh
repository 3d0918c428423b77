#!/bin/bash
# N-GPU call (N = $2, default 4): multi-GPU parity incl. world=4, weak / strong / PML scaling points.
out=gpurun_out/${1:-r01h}; N=${2:-4}; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > $out/pytest_mgpu.log 2>&1
tail -12 $out/pytest_mgpu.log | cut -c1-300
timeout 600 $TR --master-port 29521 bench.py --gpus $N --steps 200 --warmup 10 > $out/bench_n$N.json 2> $out/bench_n$N.err
cat $out/bench_n$N.json
timeout 600 $TR --master-port 29522 bench.py --gpus $N --steps 100 --warmup 10 --workload pml --no-e2e > $out/bench_pml_n$N.json 2> $out/bench_pml_n$N.err
cat $out/bench_pml_n$N.json
timeout 600 $TR --master-port 29523 bench.py --gpus $N --steps 100 --warmup 10 --size 1024 --scaling strong --no-e2e > $out/bench_strong1024_n$N.json 2> $out/bench_strong1024_n$N.err
cat $out/bench_strong1024_n$N.json
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu --no-e2e > $out/bench_n1.json 2> $out/bench_n1.err
cat $out/bench_n1.json
