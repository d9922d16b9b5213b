#!/bin/bash
out=gpurun_out/r2san; mkdir -p $out
timeout 700 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests -m gpu -q -x -k "not stress and not two_sessions and not full_size" > $out/memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed" $out/memcheck.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests -m gpu -q -x -k "test_golden_vectors or test_resize_shapes or test_overlay_vs_oracle or test_prepared_batch or test_depth16 or test_text_run_views or test_composite_sources or test_dense_overlay or test_composite_resize or test_nv12" > $out/racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed|hazard" $out/racecheck.log | tail -5
