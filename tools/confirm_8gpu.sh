# short 8-GPU confirmation after the thin-boundary-chunk change: 1024^3 strong scaling at 8 and 1, weak scaling at 8
out=gpurun_out/${TAG:-r02h}; mkdir -p $out
nvidia-smi -L > $out/gpu.txt 2>&1
S="--steps 60 --warmup 6 --reps 3 --no-e2e --no-cpu --no-verify --zero-init --size 1024 --scaling strong"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29801 bench.py --gpus 8 $S --timeline > $out/bench_strong1024_n8.json 2> $out/bench_strong1024_n8.err; cut -c1-250 $out/bench_strong1024_n8.json
mkdir -p $out/timeline_strong_n8; mv gpurun_out/timeline_n8_rank*.json $out/timeline_strong_n8/ 2>/dev/null
timeout 300 python bench.py --steps 30 --warmup 4 --reps 3 --no-e2e --no-cpu --no-verify --zero-init --size 1024 --scaling strong > $out/bench_strong1024_n1.json 2> $out/bench_strong1024_n1.err; cut -c1-250 $out/bench_strong1024_n1.json
W="--steps 200 --warmup 10 --reps 3 --no-e2e --no-cpu --no-verify"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29802 bench.py --gpus 8 $W > $out/bench_n8.json 2> $out/bench_n8.err; cut -c1-250 $out/bench_n8.json
timeout 300 python bench.py $W > $out/bench_n1.json 2> $out/bench_n1.err; cut -c1-250 $out/bench_n1.json
