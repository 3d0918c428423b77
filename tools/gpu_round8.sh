#!/bin/bash
out=gpurun_out/${1:-r01n}; mkdir -p $out
( time timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > $out/pytest_mgpu.log 2>&1
tail -4 $out/pytest_mgpu.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 200 --warmup 10 > $out/bench_n2.json 2> $out/bench_n2.err
cat $out/bench_n2.json | cut -c1-300
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu --no-e2e > $out/bench_n1.json 2> $out/bench_n1.err
cat $out/bench_n1.json | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 100 --warmup 10 --size 1024 --scaling strong --no-e2e > $out/bench_strong1024_n2.json 2> $out/bench_strong.err
cat $out/bench_strong1024_n2.json | cut -c1-300
