#!/bin/bash
# full GPU suite (parity + dist + device KSP) and the fmt_split sweep
cd "$(dirname "$0")/../.." ; mkdir -p gpurun_out
timeout 2400 python -X faulthandler -m pytest tests -x -q -m gpu --timeout 180 > gpurun_out/r8_pytest_gpu.log 2>&1; rc=$?; echo "pytest gpu rc=$rc"; tail -8 gpurun_out/r8_pytest_gpu.log
B="python bench.py --size 4096 --steps 30 --warmup 5 --no-cpu-baseline"
for sp in 8 12 20; do
timeout 600 $B --no-parity --opt fmt_split=$sp > gpurun_out/r8_b4096_split$sp.json 2> gpurun_out/r8_b4096_split$sp.log
python -c "import json;d=json.load(open('gpurun_out/r8_b4096_split$sp.json'));print('fmt_split=$sp', d['ms_per_step'], d['roofline']['frac'])"
done
