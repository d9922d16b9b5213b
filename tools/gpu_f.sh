#!/bin/bash
out=gpurun_out/r2m; mkdir -p $out
(timeout 240 python -m pytest tests -m gpu -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log)
tail -4 $out/pytest.log
for inl in 0 1; do
NES_NO_INLINE=$inl timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > $out/bench_inl$inl.json 2> $out/bench_inl$inl.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2m/bench_inl$inl.json"))
print("NO_INLINE=$inl", {k:d[k] for k in ("value","verified","single_frame_launch_fps","single_frame_api_fps","single_frame_api_host_us","p50_frame_latency_ms","p50_frame_latency_ms_2_bands") if k in d}, "e2e", round(d["e2e"]["value"]), d["e2e"]["frac_of_copy_ceiling"])
PY
done
