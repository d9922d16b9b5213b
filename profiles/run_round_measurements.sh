#!/bin/bash
# Everything the round's measured claims come from, in one gpurun call:
#   gpurun --timeout 1500 -- 'bash profiles/run_round_measurements.sh v4'
# writes gpurun_out/<tag>/; profiles/collect.sh copies the summaries into profiles/.
tag=${1:-vX}; out=gpurun_out/$tag; mkdir -p $out
(timeout 600 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log)
python bench.py > $out/bench_c2.json 2> $out/bench_c2.err
python bench.py --workload 4k_rgb24 > $out/bench_4k.json 2> $out/bench_4k.err
python bench.py --workload c3_7680x2160_sbs --steps 300 > $out/bench_c3.json 2> $out/bench_c3.err
python bench.py --workload c4_1080p_sessions --steps 300 --no-cpu-baseline > $out/bench_c4.json 2> $out/bench_c4.err
python bench.py --workload c5_4k_4src_to_1440p --steps 100 > $out/bench_c5.json 2> $out/bench_c5.err
python bench.py --impl reference --steps 10 --warmup 3 > $out/bench_ref_c2.json 2> $out/bench_ref_c2.err
python bench.py --impl reference --workload 4k_rgb24 --steps 10 --warmup 3 > $out/bench_ref_4k.json 2> $out/bench_ref_4k.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_c2.csv python bench.py --steps 3 --warmup 3 --warmup-seconds 0 --no-cpu-baseline > $out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_frame_strips -s 4 -c 1 -o $out/strips_c2 -f python bench.py --steps 3 --warmup 3 --warmup-seconds 0 --no-cpu-baseline --no-e2e > $out/ncu_full_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_frame_strips -s 4 -c 1 -o $out/strips_4k -f python bench.py --workload 4k_rgb24 --steps 3 --warmup 3 --warmup-seconds 0 --no-cpu-baseline --no-e2e > $out/ncu_full_4k.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_resize -s 2 -c 1 -o $out/resize_c5 -f python bench.py --workload c5_4k_4src_to_1440p --steps 3 --warmup 3 --warmup-seconds 0 --no-cpu-baseline --no-e2e > $out/ncu_full_c5.log 2>&1
tail -3 $out/pytest.log
for f in c2 4k c3 c4 c5 ref_c2 ref_4k; do cut -c1-160 $out/bench_$f.json; done
