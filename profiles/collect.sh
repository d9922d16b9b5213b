#!/bin/bash
# Copy the summaries of gpurun_out/<tag>/ (written by run_round_measurements.sh) into profiles/ as <tag>_*.
tag=${1:-r02x}; src=gpurun_out/$tag; pre=profiles/${tag}
for w in default ref_c2; do [ -s $src/bench_$w.json ] && cp $src/bench_$w.json ${pre}_bench_$w.json; done
for r in strips_c2 strips_4k strips_c4 resize_c5 resize_tiles_c5; do [ -s $src/$r.ncu-rep ] && python profiles/summarize_ncu.py $src/$r.ncu-rep > ${pre}_${r}_ncu_summary.csv; done
[ -s $src/launches_c2.csv ] && cp $src/launches_c2.csv ${pre}_launches_c2.csv
tail -3 $src/pytest.log > ${pre}_pytest_gpu.txt
