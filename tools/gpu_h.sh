#!/bin/bash
out=gpurun_out/r2n; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resize_strips -s 40 -c 1 -o $out/resize_c5_v2 -f python tools/diag_trace.py --workload c5_4k_4src_to_1440p --frames 0 --reps 3 > $out/ncu_full_c5_v2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 12 --csv --log-file $out/launches_c5_v2.csv python tools/diag_trace.py --workload c5_4k_4src_to_1440p --frames 0 --reps 3 > /dev/null 2>&1
tail -14 $out/launches_c5_v2.csv | cut -c1-300
