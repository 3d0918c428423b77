#!/bin/bash
# 2-GPU call: multi-GPU parity tests, stream-structure probe, 2-GPU bench; then single-GPU e2e breakdown and kc sweep.
tag=${1:-r01c}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi -L > $out/gpu.txt
( time timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > $out/pytest_mgpu.log 2>&1
tail -3 $out/pytest_mgpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_probe.py > $out/mgpu_probe.log 2>&1
grep variant $out/mgpu_probe.log
FDTD_B200_MGPU_DEBUG=8 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/mgpu_probe.py > $out/mgpu_probe_noprio.log 2>&1
grep variant $out/mgpu_probe_noprio.log | head -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 200 --warmup 10 > $out/bench_n2.json 2> $out/bench_n2.err
cat $out/bench_n2.json
timeout 300 python tools/e2e_breakdown.py > $out/e2e_breakdown.json 2> $out/e2e_breakdown.err
cat $out/e2e_breakdown.json
timeout 600 python tools/sweep.py --no-sweeps --t2 0 --kc 0 64 86 128 171 --steps 40 > $out/kc_sweep.jsonl 2>&1
cat $out/kc_sweep.jsonl
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu > $out/bench_n1.json 2> $out/bench_n1.err
cat $out/bench_n1.json
