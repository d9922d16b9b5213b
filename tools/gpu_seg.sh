#!/bin/bash
(timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -2)
for w in c2_1080p_2src_composite c4_1080p_sessions 4k_rgb24 c3_7680x2160_sbs; do timeout 100 python tools/diag_trace.py --workload $w --frames 0 --reps 50 2>&1 | grep device | sed "s/^/$w /"; done
timeout 100 python tools/diag_trace.py --workload c2_1080p_2src_composite --frames 1 --reps 200 2>&1 | grep device | sed "s/^/c2 1 frame /"
timeout 100 python tools/diag_trace.py --workload c2_1080p_2src_composite --frames 0 --reps 50 --mode convert 2>&1 | grep device | sed "s/^/c2 convert-mode /"
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench c2', round(d['value']), d['roofline']['frac'], 'e2e', round(d['e2e']['value']), d['e2e']['frac_of_copy_ceiling'], 'p50', d['p50_frame_latency_ms'], 'single', d.get('single_frame_launch_fps'), d.get('single_frame_api_fps'), d['verified']); print({k:(round(v['value']), v['roofline']['frac'], round(v['e2e']['value']), v['verified']) for k,v in d['workloads'].items()})"
