#!/bin/bash
out=gpurun_out/r2n; mkdir -p $out
echo "nproc $(nproc) cpu.max $(cat /sys/fs/cgroup/cpu.max 2>/dev/null) load $(cat /proc/loadavg)"
for cfg in "0 0" "1 0" "0 0"; do set -- $cfg
export NES_NO_INLINE=$1
timeout 300 python bench.py --workload c4_1080p_sessions --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $out/bench_c4_$1$2.json 2> $out/bench_c4_$1$2.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("$out/bench_c4_$1$2.json"))
print("NO_INLINE=$1", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["sessions"]["value"], d["sessions"]["mux"])
PY
echo "load $(cat /proc/loadavg)"; grep -E "nr_throttled|throttled_usec" /sys/fs/cgroup/cpu.stat 2>/dev/null | tr '\n' ' '; echo
done
