#!/bin/bash
# first GPU contact of the warp-tile kernel: smoke, memcheck on the smoke, parity suite, A/B bench
cd "$(dirname "$0")/../.." ; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -3 gpurun_out/r2_smoke.log
[ $rc -ne 0 ] && exit 1
timeout 900 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitize.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r2_sanitize.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_pytest_parity.log 2>&1; echo "pytest parity rc=$?"; tail -5 gpurun_out/r2_pytest_parity.log
for cfg in "kernel=2" "kernel=2 wt_stages=4" "kernel=1"; do
  tag=$(echo $cfg | tr ' =' '__')
  opts=""; for o in $cfg; do opts="$opts --opt $o"; done
  timeout 900 python bench.py --size 2048 --steps 30 --warmup 5 --no-cpu-baseline $opts --dump-ops gpurun_out/r2_ops_2048_$tag.csv > gpurun_out/r2_b2048_$tag.json 2> gpurun_out/r2_b2048_$tag.log; echo "bench 2048 $cfg rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2_b2048_$tag.json'));print('$cfg', d['ms_per_step'], d['roofline']['cycle_achieved'], d['launches_per_cycle'])"
done
