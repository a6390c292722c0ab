#!/bin/bash
# early slot re-issue: parity + 4096^2 bench (stages 2 vs 3) + per-op table
cd "$(dirname "$0")/../.." ; mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 120 > gpurun_out/r4_pytest_parity.log 2>&1; rc=$?; echo "pytest parity rc=$rc"; tail -5 gpurun_out/r4_pytest_parity.log
[ $rc -ne 0 ] && exit 1
B="python bench.py --size 4096 --steps 30 --warmup 5 --no-cpu-baseline"
timeout 1200 $B --dump-ops gpurun_out/r4_ops_4096.csv --compare-opt wt_stages=3 --compare-opt wt_stages=2,dense_rows=2048 \
   > gpurun_out/r4_b4096.json 2> gpurun_out/r4_b4096.log; echo "bench 4096 rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r4_b4096.json'));print(d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['launches_per_cycle'], d['parity']['rel_l2'], d['compare_opt_ms'], d['e2e']['ms_per_step'])"
tail -3 gpurun_out/r4_b4096.log
