#!/bin/bash
# time config 5 (dense / none / 1 source) for every library variant under build_variants/ plus the in-tree one
for so in "" $(ls build_variants/*.so 2>/dev/null); do
  export NES_GPU_LIB=$so; [ -z "$so" ] && unset NES_GPU_LIB
  echo "== ${so:-in-tree}"
  (timeout 100 python -m pytest tests -m gpu -q -x -k "config5 or c5 or resize_golden" 2>&1 | tail -1)
  for t in dense none; do
    timeout 100 python tools/diag_trace.py --workload c5_4k_4src_to_1440p --frames 0 --reps 30 --text $t 2>&1 | grep "device" | sed "s/^/$t /"
  done
  timeout 100 python tools/diag_trace.py --workload c5_4k_4src_to_1440p --frames 0 --reps 30 --nsrc 1 --text none 2>&1 | grep device | sed "s/^/1src /"
done
