# 2-GPU probe of the strong-scaling regime (every rank owns the 128-plane slab of the 1024^3 / 8-GPU configuration) + parity
out=gpurun_out/${TAG:-r02g}; mkdir -p $out
( time timeout 900 python -m pytest tests -m gpu -q -rs --deselect "tests/test_multi_gpu.py::test_zslab_ring_bit_exact[2-nccl]" --deselect "tests/test_multi_gpu.py::test_zslab_ring_bit_exact[2-peer-3stream]" ) > $out/pytest_gpu_2gpubox.log 2>&1; tail -6 $out/pytest_gpu_2gpubox.log | cut -c1-200
S="--steps 60 --warmup 6 --reps 3 --no-e2e --no-cpu --no-verify --zero-init --size 1024 --planes 256 --scaling strong"
timeout 300 python bench.py $S > $out/bench_strong_1024x1024x256_n1.json 2> $out/bench_strong_n1.err; cut -c1-250 $out/bench_strong_1024x1024x256_n1.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 $S --timeline > $out/bench_strong_1024x1024x256_n2.json 2> $out/bench_strong_n2.err; cut -c1-250 $out/bench_strong_1024x1024x256_n2.json
mkdir -p $out/timeline_strong_n2; mv gpurun_out/timeline_n2_rank*.json $out/timeline_strong_n2/ 2>/dev/null
FDTD_B200_HALO_IN_KERNEL=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 $S > $out/bench_strong_1024x1024x256_n2_3stream.json 2> $out/bench_strong_n2_3s.err; cut -c1-250 $out/bench_strong_1024x1024x256_n2_3stream.json
W="--steps 200 --warmup 10 --reps 3 --no-e2e --no-cpu --no-verify"
timeout 300 python bench.py $W > $out/bench_n1.json 2> $out/bench_n1.err; cut -c1-250 $out/bench_n1.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29703 bench.py --gpus 2 $W > $out/bench_n2.json 2> $out/bench_n2.err; cut -c1-250 $out/bench_n2.json
FDTD_B200_HALO_IN_KERNEL=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29704 bench.py --gpus 2 $W > $out/bench_n2_3stream.json 2> $out/bench_n2_3s.err; cut -c1-250 $out/bench_n2_3stream.json
