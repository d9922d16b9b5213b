#!/bin/bash
# usage: gpu_n.sh N [workload]  -- the driver's scaling launch for N GPUs
N=$1; W=${2:-c2_1080p_2src_composite}
out=gpurun_out/r2scale; mkdir -p $out
if [ "$N" = "1" ]; then
  timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --workload $W --no-extra > $out/bench_n${N}_$W.json 2> $out/bench_n${N}_$W.err
else
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 20 --warmup 5 --workload $W > $out/bench_n${N}_$W.json 2> $out/bench_n${N}_$W.err
fi
echo "rc=$?"; tail -2 $out/bench_n${N}_$W.err
python - <<PY
import json
d=json.load(open("$out/bench_n${N}_$W.json"))
print("N", d["n_gpus"], "value", round(d["value"]), "frac", d["roofline"]["frac"], "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "ceiling", round(d["e2e"]["pcie_ceiling"]["value"] or 0), "frac_ceiling", d["e2e"].get("frac_of_copy_ceiling"), "verified", d["verified"], "sessions", (d.get("sessions") or {}).get("value"), (d.get("sessions") or {}).get("mux"))
PY
