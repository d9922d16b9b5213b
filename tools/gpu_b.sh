#!/bin/bash
out=gpurun_out/r2b; mkdir -p $out
export NES_GPU_LIB=$PWD/ngp-encode-server_b200/libnes_gpu_trace.so
for cfg in "c2_1080p_2src_composite 16" "4k_rgb24 8" "c4_1080p_sessions 29"; do
  set -- $cfg
  timeout 300 python tools/diag_trace.py --workload $1 --frames $2 2>&1 | tail -2
done
cp gpurun_out/trace_*.json $out/ 2>/dev/null
