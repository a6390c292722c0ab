#!/bin/bash
# thin-warp engine: parity + 4096^2 bench (engines 2 / 1 / 0) + per-op table
cd "$(dirname "$0")/../.." ; mkdir -p gpurun_out
timeout 1200 python -X faulthandler -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 120 > gpurun_out/r6_pytest_parity.log 2>&1; rc=$?; echo "pytest parity rc=$rc"; tail -5 gpurun_out/r6_pytest_parity.log
[ $rc -ne 0 ] && exit 1
B="python bench.py --size 4096 --steps 30 --warmup 5 --no-cpu-baseline"
timeout 1200 $B --dump-ops gpurun_out/r6_ops_4096_e2.csv --compare-opt engine=1 --compare-opt engine=0 --compare-opt engine=2,ctas_per_sm=4 --compare-opt ctas_per_sm=0,pdl=0 \
   > gpurun_out/r6_b4096_e2.json 2> gpurun_out/r6_b4096_e2.log; echo "bench 4096 rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r6_b4096_e2.json'));print(d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['launches_per_cycle'], d['parity']['rel_l2'], d['compare_opt_ms'], d['e2e']['ms_per_step'])"
tail -5 gpurun_out/r6_b4096_e2.log
