# Re-check the run-time tunables of the T2 pass after a kernel change (one box, alternating): L2 promotion of the tensor maps, store policy
out=gpurun_out/${TAG:-tune}; mkdir -p $out
B="--steps 200 --warmup 10 --reps 3 --no-e2e --no-cpu --no-verify"
for cfg in "default" "FDTD_B200_TMA_L2=128" "FDTD_B200_TMA_L2=64" "FDTD_B200_ST_CS=0" "default"; do
  if [ "$cfg" = default ]; then timeout 120 python bench.py $B > $out/tmp.json 2>> $out/err.log; else env $cfg timeout 120 python bench.py $B > $out/tmp.json 2>> $out/err.log; fi
  python - "$cfg" $out/tmp.json <<'PY' | tee -a $out/tune.jsonl
import json,sys
d=json.loads(open(sys.argv[2]).read())
print(json.dumps({"cfg":sys.argv[1],"value":round(d["value"],2),"rep_ms":[round(x,1) for x in d["rep_ms"]],"sm_mhz":d["clocks"].get("sm_mhz")}))
PY
done
