#!/bin/bash
# 2-GPU call: multi-GPU parity (incl. the PML two-step pass on slabs), PML bench at N=1 and N=2.
out=gpurun_out/${1:-r01g}; mkdir -p $out
( time timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > $out/pytest_mgpu.log 2>&1
tail -15 $out/pytest_mgpu.log | cut -c1-300
timeout 600 python bench.py --workload pml --steps 100 --warmup 10 --no-cpu > $out/bench_pml_n1.json 2> $out/bench_pml_n1.err
cat $out/bench_pml_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload pml --steps 100 --warmup 10 > $out/bench_pml_n2.json 2> $out/bench_pml_n2.err
cat $out/bench_pml_n2.json
