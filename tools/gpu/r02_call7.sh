#!/bin/bash
# per-operator format (chunk for short rows, row-aligned + direct engine otherwise): parity + bench + per-op
cd "$(dirname "$0")/../.." ; mkdir -p gpurun_out
timeout 1200 python -X faulthandler -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 120 > gpurun_out/r7_pytest_parity.log 2>&1; rc=$?; echo "pytest parity rc=$rc"; tail -5 gpurun_out/r7_pytest_parity.log
[ $rc -ne 0 ] && exit 1
B="python bench.py --size 4096 --steps 30 --warmup 5 --no-cpu-baseline"
timeout 1200 $B --dump-ops gpurun_out/r7_ops_4096.csv --compare-opt engine=2 --compare-opt engine=0 --compare-opt engine=1,pdl=0 \
   > gpurun_out/r7_b4096.json 2> gpurun_out/r7_b4096.log; echo "bench 4096 rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r7_b4096.json'));print(d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['launches_per_cycle'], d['parity']['rel_l2'], d['compare_opt_ms'], d['e2e']['ms_per_step'])"
tail -4 gpurun_out/r7_b4096.log
for sp in 3 8; do
timeout 600 $B --no-parity --opt fmt_split=$sp > gpurun_out/r7_b4096_split$sp.json 2> gpurun_out/r7_b4096_split$sp.log
python -c "import json;d=json.load(open('gpurun_out/r7_b4096_split$sp.json'));print('fmt_split=$sp', d['ms_per_step'], d['roofline']['frac'])"
done
