#!/bin/bash
out=gpurun_out/r02d; mkdir -p $out
for w in "c2_1080p_2src_composite strips_c2 k_frame_strips" "c5_4k_4src_to_1440p resize_c5 k_resize_strips"; do
  set -- $w
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s 40 -c 1 -o $out/$2 -f python tools/diag_trace.py --workload $1 --frames 0 --reps 3 > $out/ncu_full_$2.log 2>&1
done
ls -la $out
