#!/bin/bash
out=gpurun_out/r2n; mkdir -p $out
(timeout 150 python -m pytest tests -m gpu -q -x -k "resize or config or golden or c5" > $out/pytest_rz.log 2>&1; echo "pytest rc=$?" >> $out/pytest_rz.log)
tail -2 $out/pytest_rz.log
for t in dense none; do
  timeout 120 python tools/diag_trace.py --workload c5_4k_4src_to_1440p --frames 0 --reps 30 --text $t 2>&1 | grep "device"
done
timeout 120 python tools/diag_trace.py --workload c5_4k_4src_to_1440p --frames 0 --reps 30 --nsrc 1 --text none 2>&1 | grep device
