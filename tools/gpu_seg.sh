#!/bin/bash
for i in 1 2; do
timeout 100 python tools/diag_trace.py --workload c3_7680x2160_sbs --frames 1 --reps 200 --mode convert 2>&1 | grep device | sed "s/^/c3 1 frame convert /"
timeout 100 python tools/diag_trace.py --workload c3_7680x2160_sbs --frames 4 --reps 100 --mode convert 2>&1 | grep device | sed "s/^/c3 4 frames convert /"
timeout 100 python tools/diag_trace.py --workload c3_7680x2160_sbs --frames 1 --reps 200 2>&1 | grep device | sed "s/^/c3 1 frame prepared /"
done
timeout 300 python bench.py --workload c3_7680x2160_sbs --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench c3', round(d['value']), d['roofline']['frac'], 'single', d.get('single_frame_launch_fps'), d.get('single_frame_api_fps'), d.get('single_frame_api_host_us'))"
