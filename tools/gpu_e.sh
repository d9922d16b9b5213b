#!/bin/bash
W=c5_4k_4src_to_1440p
for dw in 128 64 48; do
echo "MAXDW $dw"
NES_RZ_MAXDW=$dw timeout 100 python tools/diag_trace.py --workload $W --frames 3 2>&1 | tail -1
NES_RZ_MAXDW=$dw timeout 100 python tools/diag_trace.py --workload $W --frames 3 --text none 2>&1 | tail -1
done
