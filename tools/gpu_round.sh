#!/bin/bash
# One entry point for every GPU-box run of a round.  EVERY command runs under `timeout` (a hung NCCL experiment once
# ate the rest of a round's GPU budget).  Usage (from the repo root, normally through gpurun):
#
#     bash tools/gpu_round.sh <tag> <stage> [<stage> ...]          output under gpurun_out/<tag>/
#
# 1-GPU stages
#   tests        GPU parity suite (pytest -m gpu) + __graft_entry__.smoke()
#   bench        bench.py (ours) and bench.py --impl reference
#   table        tools/config_table.py (BASELINE configs on one GPU)
#   launches     ncu launch list (gpu__time_duration.sum) of the bench command
#   ncu          ncu --set full of the T2 pass, fp64 and fp32
#   pml          PML launch list with DRAM bytes + ncu --set full of the sweeps + PML bench line
#   sanitize     compute-sanitizer memcheck / racecheck / synccheck over tools/sanitize_driver.py
#   e2e          tools/e2e_breakdown.py
#   kc           DRAM traffic / duration of the T2 pass vs chunk length (512^3 and 1024^3)
# N-GPU stages (set NGPU=2|4|8 and call gpurun --gpus $NGPU)
#   mtests       tests/test_multi_gpu.py (every transport; one-process ring)
#   mcoarray     tests/test_cpp_api_gpu.py (coarray-style program with $NGPU images, C++ multi-GPU class)
#   mbench       1-GPU bench on the same box, weak-scaling bench at $NGPU with the per-pass timeline, PML weak, 1024^3 strong
#   mdriver      the driver's own command line at $NGPU
#   mab          the other two halo transports at $NGPU, with timelines
#   mprobe       tools/mgpu_probe.py (stream-structure variants)
tag=${1:?tag}; shift
out=gpurun_out/$tag; mkdir -p $out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L > $out/gpu.txt 2>&1; nproc >> $out/gpu.txt

ncu_rows() {   # csv -> one json line per launch: duration + dram bytes
python - "$@" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
d = {}
for r in rows:
    d.setdefault(r[0], {"kernel": r[4] if len(r) > 4 else ""})[r[-3]] = float(r[-1].replace(",", ""))
for k, v in d.items():
    rd, wr, t = v.get("dram__bytes_read.sum", 0), v.get("dram__bytes_write.sum", 0), v.get("gpu__time_duration.sum", 0)
    print('{"id": %s, "ms": %.3f, "read_GB": %.2f, "write_GB": %.2f, "GBs": %.0f, "tag": "%s"}' % (k, t / 1e6, rd / 1e9, wr / 1e9, (rd + wr) / max(t, 1), " ".join(sys.argv[2:])))
PY
}

for stage in "$@"; do
case $stage in
tests)
  ( time timeout 900 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1; tail -3 $out/pytest_gpu.log
  timeout 300 python __graft_entry__.py --smoke > $out/smoke.log 2>&1; tail -2 $out/smoke.log ;;
bench)
  timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench_n1_driver_args.json 2> $out/bench_n1_driver_args.err; cut -c1-400 $out/bench_n1_driver_args.json
  timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu > $out/bench_n1.json 2> $out/bench_n1.err; cut -c1-300 $out/bench_n1.json
  timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref.json 2> $out/bench_ref.err; cut -c1-200 $out/bench_ref.json ;;
table)
  timeout 900 python tools/config_table.py > $out/config_table.jsonl 2> $out/config_table.err; cut -c1-220 $out/config_table.jsonl ;;
launches)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
      python bench.py --steps 20 --warmup 5 --reps 2 --no-cpu --no-verify > $out/bench_under_ncu.log 2>&1 ;;
ncu)
  for dt in f64 f32 f32a; do
    extra=""; d=$dt; if [ $dt = f32a ]; then d=f32; extra="--f32-arith"; fi
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_BE_T2 -s 4 -c 1 -f -o $out/t2_${dt}_full \
        python bench.py --dtype $d $extra --steps 12 --warmup 3 --reps 1 --no-cpu --no-e2e --no-verify > $out/ncu_$dt.log 2>&1
  done ;;
pml)
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file $out/launches_pml.csv \
      python bench.py --workload pml --steps 4 --warmup 3 --no-cpu --no-e2e > $out/bench_pml_under_ncu.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_ -s 12 -c 4 -f -o $out/pml_full \
      python bench.py --workload pml --steps 4 --warmup 3 --no-cpu --no-e2e > $out/ncu_pml_full.log 2>&1
  timeout 600 python bench.py --workload pml --steps 100 --warmup 10 --no-cpu > $out/bench_pml.json 2> $out/bench_pml.err; cut -c1-300 $out/bench_pml.json ;;
sanitize)   # compute-sanitizer memcheck + racecheck (shared-memory hazards of the single-buffered row exchanges) + synccheck
  for tool in memcheck racecheck synccheck; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_driver.py > $out/sanitizer_$tool.log 2>&1
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|all cases|Error|hazard" $out/sanitizer_$tool.log | head -8
  done ;;
e2e)
  timeout 300 python tools/e2e_breakdown.py --steps 60 > $out/e2e_breakdown.json 2> $out/e2e_breakdown.err; cat $out/e2e_breakdown.json ;;
kc)
  for cfg in "512 32" "512 64" "512 171" "1024 32" "1024 64" "1024 128" "1024 512"; do
    read n kc <<< "$cfg"
    timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:fused_BE_T2 -s 2 -c 1 --csv \
        --log-file $out/kc_${n}_$kc.csv python tools/sweep.py --n $n --no-sweeps --t2 0 --kc $kc --steps 6 > /dev/null 2>&1
    ncu_rows $out/kc_${n}_$kc.csv n=$n kc=$kc
  done | tee $out/kc_traffic.jsonl ;;
mtests)
  ( time timeout 1500 python -m pytest tests/test_multi_gpu.py -m gpu -q -rs -v ${MTESTS_K:+-k "$MTESTS_K"} ) > $out/pytest_mgpu_n$N.log 2>&1; tail -14 $out/pytest_mgpu_n$N.log | cut -c1-200 ;;
mcoarray)   # the coarray-style program with $N images (cpp/coarray, f4) + the C++ multi-GPU class tests
  ( timeout 600 python -m pytest tests/test_cpp_api_gpu.py -m gpu -q -rs -v ) > $out/pytest_cpp_n$N.log 2>&1; tail -24 $out/pytest_cpp_n$N.log | cut -c1-200
  ( timeout 300 cpp/bin/fdtd_coarray_b200 --images $N 64 64 $((32 * N)) 50 ) > $out/coarray_${N}_images.log 2>&1; head -8 $out/coarray_${N}_images.log ;;
mbench)     # weak scaling (the driver's contract) with the per-pass timeline, same-box N = 1, PML weak, 1024^3 strong
  timeout 600 python bench.py --steps 200 --warmup 10 --reps 3 --no-cpu --no-e2e --no-verify > $out/bench_n1_samebox.json 2> $out/bench_n1_samebox.err; cut -c1-300 $out/bench_n1_samebox.json
  timeout 600 $TR --master-port 29521 bench.py --gpus $N --steps 200 --warmup 10 --reps 3 --no-e2e --timeline > $out/bench_n$N.json 2> $out/bench_n$N.err; cut -c1-300 $out/bench_n$N.json
  mkdir -p $out/timeline_n$N; mv gpurun_out/timeline_n${N}_rank*.json $out/timeline_n$N/ 2>/dev/null
  timeout 600 $TR --master-port 29522 bench.py --gpus $N --steps 100 --warmup 10 --reps 3 --workload pml --no-e2e > $out/bench_pml_n$N.json 2> $out/bench_pml_n$N.err; cut -c1-300 $out/bench_pml_n$N.json
  timeout 600 $TR --master-port 29523 bench.py --gpus $N --steps 100 --warmup 10 --reps 3 --size 1024 --scaling strong --no-e2e > $out/bench_strong1024_n$N.json 2> $out/bench_strong1024_n$N.err; cut -c1-300 $out/bench_strong1024_n$N.json ;;
mdriver)    # exactly what the driver runs at N > 1 (short window, e2e leg included)
  timeout 600 $TR --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > $out/bench_n${N}_driver_args.json 2> $out/bench_n${N}_driver_args.err; cut -c1-300 $out/bench_n${N}_driver_args.json ;;
m8)         # the one 8-GPU session of a round: same-box series at N = 1, 4, 8 for the three configurations + parity at 8 ranks
  ( time timeout 600 python -m pytest "tests/test_multi_gpu.py::test_zslab_ring_bit_exact[8-peer-inkernel]" -m gpu -q -rs ) > $out/pytest_mgpu_n8.log 2>&1; tail -5 $out/pytest_mgpu_n8.log | cut -c1-200
  B="--steps 200 --warmup 10 --reps 3 --no-e2e --no-cpu"
  timeout 300 python bench.py $B --no-verify > $out/bench_n1.json 2> $out/bench_n1.err; cut -c1-200 $out/bench_n1.json
  for n in 8 4; do
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n $B --timeline > $out/bench_n$n.json 2> $out/bench_n$n.err; cut -c1-200 $out/bench_n$n.json
    mkdir -p $out/timeline_n$n; mv gpurun_out/timeline_n${n}_rank*.json $out/timeline_n$n/ 2>/dev/null
  done
  FDTD_B200_TRANSPORT=nccl timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 $B --timeline > $out/bench_n8_nccl.json 2> $out/bench_n8_nccl.err; cut -c1-200 $out/bench_n8_nccl.json
  mkdir -p $out/timeline_n8_nccl; mv gpurun_out/timeline_n8_rank*.json $out/timeline_n8_nccl/ 2>/dev/null
  P="--steps 100 --warmup 10 --reps 3 --no-e2e --no-cpu --workload pml"
  timeout 300 python bench.py $P > $out/bench_pml_n1.json 2> $out/bench_pml_n1.err; cut -c1-200 $out/bench_pml_n1.json
  for n in 8 4; do
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n bench.py --gpus $n $P > $out/bench_pml_n$n.json 2> $out/bench_pml_n$n.err; cut -c1-200 $out/bench_pml_n$n.json
  done
  S="--steps 60 --warmup 6 --reps 3 --no-e2e --no-cpu --no-verify --zero-init --size 1024 --scaling strong"
  for n in 8 4; do
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2963$n bench.py --gpus $n $S > $out/bench_strong1024_n$n.json 2> $out/bench_strong1024_n$n.err; cut -c1-200 $out/bench_strong1024_n$n.json
  done
  timeout 500 python bench.py --steps 30 --warmup 4 --reps 3 --no-e2e --no-cpu --no-verify --zero-init --size 1024 --scaling strong > $out/bench_strong1024_n1.json 2> $out/bench_strong1024_n1.err; cut -c1-200 $out/bench_strong1024_n1.json
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 8 --steps 20 --warmup 5 > $out/bench_n8_driver_args.json 2> $out/bench_n8_driver_args.err; cut -c1-200 $out/bench_n8_driver_args.json ;;
mab)        # transports side by side: copy engines + halo wait in the kernel (default) / + three-stream boundary launches / NCCL
  FDTD_B200_HALO_IN_KERNEL=0 timeout 600 $TR --master-port 29524 bench.py --gpus $N --steps 200 --warmup 10 --reps 3 --no-e2e --timeline > $out/bench_n${N}_peer3stream.json 2> $out/bench_n${N}_peer3stream.err; cut -c1-300 $out/bench_n${N}_peer3stream.json
  mkdir -p $out/timeline_n${N}_peer3stream; mv gpurun_out/timeline_n${N}_rank*.json $out/timeline_n${N}_peer3stream/ 2>/dev/null
  FDTD_B200_TRANSPORT=nccl timeout 600 $TR --master-port 29525 bench.py --gpus $N --steps 200 --warmup 10 --reps 3 --no-e2e --timeline > $out/bench_n${N}_nccl.json 2> $out/bench_n${N}_nccl.err; cut -c1-300 $out/bench_n${N}_nccl.json
  mkdir -p $out/timeline_n${N}_nccl; mv gpurun_out/timeline_n${N}_rank*.json $out/timeline_n${N}_nccl/ 2>/dev/null ;;
mprobe)
  timeout 600 $TR --master-port 29511 tools/mgpu_probe.py > $out/mgpu_probe.log 2>&1; grep variant $out/mgpu_probe.log ;;
*) echo "unknown stage $stage" ;;
esac
done
ls -la $out | tail -n +2 | head -40
