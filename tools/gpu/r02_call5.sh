#!/bin/bash
# direct engine + fused permutation: parity + 4096^2 bench (engine 1 vs 0) + per-op table
cd "$(dirname "$0")/../.." ; mkdir -p gpurun_out
timeout 1200 python -X faulthandler -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 120 > gpurun_out/r5_pytest_parity.log 2>&1; rc=$?; echo "pytest parity rc=$rc"; tail -5 gpurun_out/r5_pytest_parity.log
[ $rc -ne 0 ] && exit 1
B="python bench.py --size 4096 --steps 30 --warmup 5 --no-cpu-baseline"
timeout 1200 $B --dump-ops gpurun_out/r5_ops_4096_e1.csv --compare-opt engine=0 --compare-opt engine=1,ctas_per_sm=2 --compare-opt ctas_per_sm=0,pdl=0 \
   > gpurun_out/r5_b4096_e1.json 2> gpurun_out/r5_b4096_e1.log; echo "bench 4096 rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r5_b4096_e1.json'));print(d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['launches_per_cycle'], d['parity']['rel_l2'], d['compare_opt_ms'], d['e2e']['ms_per_step'])"
tail -4 gpurun_out/r5_b4096_e1.log
timeout 1200 $B --no-parity --opt fuse_perm=0 > gpurun_out/r5_b4096_nofuse.json 2> gpurun_out/r5_b4096_nofuse.log; echo "bench 4096 nofuse rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r5_b4096_nofuse.json'));print('fuse_perm=0', d['ms_per_step'], d['roofline']['frac'], d['launches_per_cycle'])"
