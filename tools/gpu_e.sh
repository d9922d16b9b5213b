#!/bin/bash
W=c5_4k_4src_to_1440p
(timeout 150 python -m pytest tests -m gpu -q -x -k "resize or golden or config_size or nv12 or composite or depth16" > /tmp/p.log 2>&1; tail -2 /tmp/p.log)
timeout 100 python tools/diag_trace.py --workload $W --frames 3 2>&1 | tail -1
timeout 100 python tools/diag_trace.py --workload $W --frames 3 --text none 2>&1 | tail -1
timeout 100 python tools/diag_trace.py --workload $W --frames 3 --text none --nsrc 1 2>&1 | tail -1
