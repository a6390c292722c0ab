#!/bin/bash
# row-aligned warp-tile format: parity + 4096^2 bench (stages 2 vs 3) + per-op table + launch list
cd "$(dirname "$0")/../.." ; mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 120 > gpurun_out/r3_pytest_parity.log 2>&1; rc=$?; echo "pytest parity rc=$rc"; tail -5 gpurun_out/r3_pytest_parity.log
[ $rc -ne 0 ] && exit 1
B="python bench.py --size 4096 --steps 30 --warmup 5 --no-cpu-baseline"
timeout 1200 $B --dump-ops gpurun_out/r3_ops_4096.csv --compare-opt wt_stages=3 --compare-opt wt_stages=2,pdl=0 \
   > gpurun_out/r3_b4096.json 2> gpurun_out/r3_b4096.log; echo "bench 4096 rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r3_b4096.json'));print(d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['launches_per_cycle'], d['parity']['rel_l2'], d['compare_opt_ms'], d['e2e']['ms_per_step'])"
tail -3 gpurun_out/r3_b4096.log
N=$(cat gpurun_out/r3_ops_4096.csv.nspmv)
P="$B --no-parity --profile-one-cycle"
$P > gpurun_out/r3_plain_profile.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_elapsed.max --clock-control none -c 400 --csv --log-file gpurun_out/r3_launches_4096.csv $P > gpurun_out/r3_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
