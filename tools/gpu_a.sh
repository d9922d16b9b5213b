#!/bin/bash
# first GPU call of round 2: the new parity tests + per-CTA timelines of k_frame_strips
out=gpurun_out/r2a; mkdir -p $out
(timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log)
tail -5 $out/pytest.log
export NES_GPU_LIB=$PWD/ngp-encode-server_b200/libnes_gpu_trace.so
for cfg in "c2_1080p_2src_composite 16" "c2_1080p_2src_composite 64" "c4_1080p_sessions 29" "4k_rgb24 8" "4k_rgb24 1" "c2_1080p_2src_composite 1"; do
  set -- $cfg
  timeout 300 python tools/diag_trace.py --workload $1 --frames $2 2>&1 | tail -2
done
cp gpurun_out/trace_*.json $out/ 2>/dev/null
