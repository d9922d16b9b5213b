#!/bin/bash
out=gpurun_out/r2j; mkdir -p $out
(timeout 240 python -m pytest tests -m gpu -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log)
tail -4 $out/pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --workload c4_1080p_sessions --no-cpu-baseline > $out/bench_c4.json 2> $out/bench_c4.err; echo "bench rc=$?"
tail -3 $out/bench_c4.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2j/bench_c4.json"))
print({k:d[k] for k in ("value","verified","single_frame_launch_fps","single_frame_api_fps","p50_frame_latency_ms") if k in d})
print("roofline", d["roofline"]["frac"]); print("e2e", d["e2e"]["value"], d["e2e"].get("frac_of_copy_ceiling")); print("sessions", d.get("sessions"))
PY
