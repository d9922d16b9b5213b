#!/bin/bash
# Everything the round's measured claims come from, in one gpurun call:
#   gpurun --timeout 1500 -- 'bash profiles/run_round_measurements.sh r02a'
# writes gpurun_out/<tag>/; profiles/collect.sh <tag> copies the summaries into profiles/.
# (every step runs under its own timeout: a hung kernel must not eat the GPU budget)
tag=${1:-r02x}; out=gpurun_out/$tag; mkdir -p $out
(timeout 300 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log)
tail -3 $out/pytest.log
# the driver's two arms, exactly as it launches them
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $out/bench_ref_c2.json 2> $out/bench_ref_c2.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/bench_default.json 2> $out/bench_default.err
# the launch list of the same command (cold-cache, serialised: shares, not absolutes) and one full capture per kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_c2.csv python bench.py --steps 2 --warmup 3 --warmup-seconds 0 --no-cpu-baseline --no-extra > $out/ncu_launches.log 2>&1
for w in "c2_1080p_2src_composite strips_c2 k_frame_strips" "4k_rgb24 strips_4k k_frame_strips" "c4_1080p_sessions strips_c4 k_frame_strips" "c5_4k_4src_to_1440p resize_c5 k_resize_strips"; do
  set -- $w
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s 40 -c 1 -o $out/$2 -f python tools/diag_trace.py --workload $1 --frames 0 --reps 3 > $out/ncu_full_$2.log 2>&1
done
NES_NO_RZ=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resize_tiles -s 40 -c 1 -o $out/resize_tiles_c5 -f python tools/diag_trace.py --workload c5_4k_4src_to_1440p --frames 0 --reps 3 > $out/ncu_full_tiles.log 2>&1
python - <<PY
import json
d=json.load(open("$out/bench_default.json"))
print("c2", round(d["value"]), d["roofline"]["frac"], "e2e", round(d["e2e"]["value"]), "verified", d["verified"])
for k,v in d.get("workloads",{}).items(): print(k, round(v.get("value",0)), v.get("roofline",{}).get("frac"), "e2e", round(v.get("e2e",{}).get("value",0)), v.get("verified"))
r=json.load(open("$out/bench_ref_c2.json")); print("ref", r["value"], r["cpu_baseline"]["cores"])
PY
ls $out
