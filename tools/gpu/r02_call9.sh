#!/bin/bash
cd "$(dirname "$0")/../.." ; mkdir -p gpurun_out
B="python bench.py --size 4096 --steps 30 --warmup 5 --no-cpu-baseline --no-parity"
timeout 1200 $B --compare-opt sv_minb=4 --compare-opt sv_minb=4,dense_rows=2048 --compare-opt sv_minb=3,dense_rows=8192 --compare-opt dense_rows=4096,engine=2 \
   > gpurun_out/r9_b4096.json 2> gpurun_out/r9_b4096.log; echo "bench 4096 rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r9_b4096.json'));print(d['ms_per_step'], d['roofline']['frac'], d['compare_opt_ms'])"
