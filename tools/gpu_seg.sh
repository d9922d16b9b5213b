#!/bin/bash
for inf in 3 4 6; do
  for w in c3_7680x2160_sbs c2_1080p_2src_composite c4_1080p_sessions; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-extra --in-flight $inf 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('in-flight $inf $w e2e', round(e['value']), 'ceil', round(e['pcie_ceiling']['value']), e['frac_of_copy_ceiling'], 'p50', d['p50_frame_latency_ms'])"
  done
done
