#!/bin/bash
out=gpurun_out/r2f; mkdir -p $out
(timeout 900 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log)
tail -8 $out/pytest.log
for cfg in "c5_4k_4src_to_1440p 3" "c5_4k_4src_to_1440p 1" "c5_4k_4src_to_1440p 6"; do
  set -- $cfg
  timeout 300 python tools/diag_trace.py --workload $1 --frames $2 --mode prepared 2>&1 | tail -2
  NES_NO_RZ=1 timeout 300 python tools/diag_trace.py --workload $1 --frames $2 --mode prepared 2>&1 | tail -2
done
