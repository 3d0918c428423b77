# Final validation of a round on one 2-GPU box: the whole GPU suite (multi-GPU matrix at 2 ranks included), smoke(), the driver's
# command lines at N = 1 and N = 2, the reference arm, the one-GPU config table, ncu captures of the final T2 kernels.
out=gpurun_out/${TAG:-final}; mkdir -p $out
nvidia-smi -L > $out/gpu.txt 2>&1; nproc >> $out/gpu.txt
( time timeout 1200 python -m pytest tests -m gpu -q -rs ) > $out/pytest_gpu.log 2>&1; tail -8 $out/pytest_gpu.log | cut -c1-200
timeout 300 python __graft_entry__.py --smoke > $out/smoke.log 2>&1; tail -2 $out/smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/bench_n1_driver_args.json 2> $out/bench_n1.err; cut -c1-300 $out/bench_n1_driver_args.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 2 --steps 20 --warmup 5 > $out/bench_n2_driver_args.json 2> $out/bench_n2.err; cut -c1-300 $out/bench_n2_driver_args.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $out/bench_reference_arm.json 2> $out/bench_ref.err; cut -c1-300 $out/bench_reference_arm.json
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu > $out/bench_n1_200.json 2> $out/bench_n1_200.err; cut -c1-200 $out/bench_n1_200.json
timeout 900 python tools/config_table.py --quick --skip-1024 > $out/config_table.jsonl 2> $out/config_table.err; cut -c1-200 $out/config_table.jsonl
for dt in f64 f32a; do
  extra=""; d=$dt; if [ $dt = f32a ]; then d=f32; extra="--f32-arith"; fi
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_BE_T2 -s 4 -c 1 -f -o $out/t2_${dt}_full \
      python bench.py --dtype $d $extra --steps 12 --warmup 3 --reps 1 --no-cpu --no-e2e --no-verify > $out/ncu_$dt.log 2>&1
done
ls -la $out | tail -n +2 | head -30
