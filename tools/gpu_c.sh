#!/bin/bash
out=gpurun_out/r2d; mkdir -p $out
export NES_GPU_LIB=$PWD/ngp-encode-server_b200/libnes_gpu_trace.so
for mode in convert prepared; do
for cfg in "c2_1080p_2src_composite 16" "c4_1080p_sessions 29" "c2_1080p_2src_composite 1" "4k_rgb24 1" "c2_1080p_2src_composite 64" "c4_1080p_sessions 64"; do
  set -- $cfg
  timeout 300 python tools/diag_trace.py --workload $1 --frames $2 --mode $mode 2>&1 | tail -2
  cp gpurun_out/trace_$1_$2.json $out/trace_$1_$2_$mode.json
done
done
