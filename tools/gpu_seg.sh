#!/bin/bash
(timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -2)
for t in dense none; do timeout 100 python tools/diag_trace.py --workload c5_4k_4src_to_1440p --frames 0 --reps 30 --text $t 2>&1 | grep device | sed "s/^/$t /"; done
timeout 100 python tools/diag_trace.py --workload c5_4k_4src_to_1440p --frames 1 --reps 30 2>&1 | grep device | sed "s/^/1 frame /"
timeout 100 python tools/diag_trace.py --workload c5_4k_4src_to_1440p --frames 0 --reps 30 --nsrc 1 --text none 2>&1 | grep device | sed "s/^/1src /"
timeout 300 python bench.py --workload c5_4k_4src_to_1440p --steps 20 --warmup 5 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench c5', round(d['value']), d['roofline']['frac'], d['roofline']['launch_us'], 'e2e', round(d['e2e']['value']), 'p50', d['p50_frame_latency_ms'], 'single', d.get('single_frame_launch_fps'), d['verified'])"
