#!/bin/bash
out=gpurun_out/r2c; mkdir -p $out
(timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log)
tail -5 $out/pytest.log
export NES_GPU_LIB=$PWD/ngp-encode-server_b200/libnes_gpu_trace.so
for pdl in 0 1; do
for cfg in "c2_1080p_2src_composite 16" "4k_rgb24 8" "c4_1080p_sessions 29" "c2_1080p_2src_composite 1" "4k_rgb24 1"; do
  set -- $cfg
  echo "NO_PDL=$pdl"
  NES_NO_PDL=$pdl timeout 300 python tools/diag_trace.py --workload $1 --frames $2 2>&1 | tail -1
  cp gpurun_out/trace_$1_$2.json $out/trace_$1_$2_nopdl$pdl.json
done
done
