#!/bin/bash
cd "$(dirname "$0")/../.." ; mkdir -p gpurun_out
free -g | head -2
(while true; do sleep 5; ps -o rss,vsz,etime,cmd -C python | tail -2 >> gpurun_out/r3_mem.log; done) &
MON=$!
timeout 400 python -X faulthandler -m pytest tests/test_gpu_parity.py -x -v -m gpu --timeout 90 -k "test_vcycle_parity" > gpurun_out/r3_diag.log 2>&1; echo "rc=$?"
kill $MON
tail -40 gpurun_out/r3_diag.log
tail -5 gpurun_out/r3_mem.log
