#!/bin/bash
out=gpurun_out/r2san; mkdir -p $out
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests -m gpu -q -x -k "test_golden_vectors or test_resize_shapes or test_overlay_vs_oracle or test_prepared_batch or test_depth16 or test_text_run_views or test_composite_sources or test_dense_overlay or test_composite_resize or test_nv12" > $out/racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed" $out/racecheck.log | tail -5
grep -E "Race reported" $out/racecheck.log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | head
for t in dense none; do timeout 100 python tools/diag_trace.py --workload c5_4k_4src_to_1440p --frames 0 --reps 30 --text $t 2>&1 | grep device; done
