#!/bin/bash
cd "$(dirname "$0")/../.." ; mkdir -p gpurun_out
B0="python bench.py --size 4096 --steps 30 --warmup 5 --no-cpu-baseline --no-parity"
timeout 600 $B0 --compare-opt sv_pf=1 --compare-opt sv_pf=1,fmt_split=8 > gpurun_out/r13_b4096.json 2> gpurun_out/r13_b4096.log
python -c "import json;d=json.load(open('gpurun_out/r13_b4096.json'));print('4096^2', d['ms_per_step'], d['roofline']['frac'], d['compare_opt_ms'])"
timeout 1500 python bench.py --workload adv_diff_fd_3d_lair --size 256 --steps 20 --warmup 3 --no-cpu-baseline --compare-opt sv_pf=1 --dump-ops gpurun_out/r13_ops_3d256.csv > gpurun_out/r13_b3d256.json 2> gpurun_out/r13_b3d256.log; echo "3d rc=$?"
grep "\[bench\]" gpurun_out/r13_b3d256.log | tail -6
python -c "import json;d=json.load(open('gpurun_out/r13_b3d256.json'));print('3D 256^3 lAIR', d['ms_per_step'], d['value'], d['roofline']['achieved'], d['roofline']['frac'], d['launches_per_cycle'], d['parity']['rel_l2'], d['compare_opt_ms'])"
