#!/bin/bash
# scaling run at N GPUs, the driver's launch line: config 2 (default) and config 4 (64 sessions sharded s % n)
N=${1:-8}; out=gpurun_out/scale; mkdir -p $out
for wl in c2_1080p_2src_composite c4_1080p_sessions; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --workload $wl --no-extra --no-cpu-baseline > $out/n${N}_$wl.json 2> $out/n${N}_$wl.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open("$out/n${N}_$wl.json"))
print("$wl N=$N value", round(d["value"]), "frac", d["roofline"]["frac"], "e2e", round(d["e2e"]["value"]), "ceil", round(d["e2e"]["pcie_ceiling"]["value"]), d["e2e"]["frac_of_copy_ceiling"], "sessions", (d.get("sessions") or {}).get("value"), (d.get("sessions") or {}).get("mux"), "verified", d["verified"])
PY
done
