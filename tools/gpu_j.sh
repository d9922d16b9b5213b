#!/bin/bash
out=gpurun_out/r2n; mkdir -p $out
timeout 300 ncu --metrics gpu__time_duration.sum,launch__grid_size,smsp__inst_executed.sum --clock-control none -s 100 -c 6 --csv --log-file $out/launches_c5_v2.csv python tools/diag_trace.py --workload c5_4k_4src_to_1440p --frames 0 --reps 3 > /dev/null 2>&1
cut -d, -f5,12- $out/launches_c5_v2.csv | tail -20 | cut -c1-250
