#!/bin/bash
# resize kernel iteration: parity of the resize / overlay / batch tests first (short timeout: a protocol bug hangs), then the
# whole suite, then config 5 timings
out=gpurun_out/r2n; mkdir -p $out
(timeout 150 python -m pytest tests -m gpu -q -x -k "resize or config or golden or c5 or overlay or text or batch or mux or nv12 or depth16" > $out/pytest_rz.log 2>&1; echo "pytest rc=$?" >> $out/pytest_rz.log)
tail -5 $out/pytest_rz.log
if grep -q "rc=0" $out/pytest_rz.log; then
  (timeout 300 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log); tail -3 $out/pytest.log
  for t in dense none; do
    timeout 120 python tools/diag_trace.py --workload c5_4k_4src_to_1440p --frames 0 --reps 30 --text $t > $out/diag_c5_$t.log 2>&1; echo "diag $t rc=$?"; tail -4 $out/diag_c5_$t.log
  done
  timeout 120 python tools/diag_trace.py --workload c5_4k_4src_to_1440p --frames 0 --reps 30 --nsrc 1 > $out/diag_c5_1src.log 2>&1; tail -2 $out/diag_c5_1src.log
  timeout 300 python bench.py --workload c5_4k_4src_to_1440p --steps 20 --warmup 5 --no-cpu-baseline --no-extra > $out/bench_c5.json 2> $out/bench_c5.err; echo "bench rc=$?"
  python - <<PY
import json
d=json.load(open("$out/bench_c5.json"))
print("c5", round(d["value"]), d["roofline"], "e2e", round(d["e2e"]["value"]), d["e2e"].get("frac_of_copy_ceiling"), "verified", d.get("verified"))
PY
fi
