#!/bin/bash
# the configurations the target is stated on: 3D FD with lAIR Z (configs[3], 256^3) and the DG matrix-free surrogate (configs[2], ~50M rows)
cd "$(dirname "$0")/../.." ; mkdir -p gpurun_out
nproc; free -g | head -2
B0="python bench.py --size 4096 --steps 30 --warmup 5 --no-cpu-baseline --no-parity"
timeout 600 $B0 --compare-opt sv_minb=5 > gpurun_out/r10_b4096.json 2> gpurun_out/r10_b4096.log
python -c "import json;d=json.load(open('gpurun_out/r10_b4096.json'));print('4096^2', d['ms_per_step'], d['roofline']['frac'], d['compare_opt_ms'])"
timeout 1500 python bench.py --workload adv_diff_fd_3d_lair --size 256 --steps 20 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/r10_ops_3d256.csv > gpurun_out/r10_b3d256.json 2> gpurun_out/r10_b3d256.log; echo "3d rc=$?"
grep "\[bench\]" gpurun_out/r10_b3d256.log | tail -6
python -c "import json;d=json.load(open('gpurun_out/r10_b3d256.json'));print('3D 256^3 lAIR', d['ms_per_step'], d['value'], d['roofline']['achieved'], d['roofline']['frac'], d['launches_per_cycle'], d['parity'], d['config']['levels'], d['details']['streamed_GB_per_cycle'])"
rm -f /dev/shm/pflare_b200_cache/adv_diff_fd_3d_lair*
timeout 1800 python bench.py --workload dg_upwind --size 4096 --steps 10 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/r10_ops_dg4096.csv > gpurun_out/r10_bdg4096.json 2> gpurun_out/r10_bdg4096.log; echo "dg rc=$?"
grep "\[bench\]" gpurun_out/r10_bdg4096.log | tail -6
python -c "import json;d=json.load(open('gpurun_out/r10_bdg4096.json'));print('DG 4096^2x3 mf', d['ms_per_step'], d['value'], d['roofline']['achieved'], d['roofline']['frac'], d['launches_per_cycle'], d['parity'], d['config']['levels'], d['details']['streamed_GB_per_cycle'])"
