#!/bin/bash
# long rows outside the tiles: parity (+ dist) and the 3D workload at 160^3
cd "$(dirname "$0")/../.." ; mkdir -p gpurun_out
timeout 2400 python -X faulthandler -m pytest tests -q -m gpu --timeout 180 > gpurun_out/r11_pytest_gpu.log 2>&1; rc=$?; echo "pytest gpu rc=$rc"; tail -6 gpurun_out/r11_pytest_gpu.log
[ $rc -ne 0 ] && exit 1
timeout 900 python bench.py --workload adv_diff_fd_3d_lair --size 160 --steps 20 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/r11_ops_3d160.csv > gpurun_out/r11_b3d160.json 2> gpurun_out/r11_b3d160.log; echo "3d rc=$?"
grep "\[bench\]" gpurun_out/r11_b3d160.log | tail -4
python -c "import json;d=json.load(open('gpurun_out/r11_b3d160.json'));print('3D 160^3 lAIR', d['ms_per_step'], d['value'], d['roofline']['achieved'], d['roofline']['frac'], d['launches_per_cycle'], d['parity']['rel_l2'])"
python - <<'PY'
import csv,collections
R=list(csv.DictReader(open('gpurun_out/r11_ops_3d160.csv')))
byop=collections.defaultdict(lambda:[0,0])
for r in R: byop[r['op']][0]+=float(r['alg_bytes']); byop[r['op']][1]+=float(r['ms'])
for k,(b,t) in sorted(byop.items(),key=lambda x:-x[1][1]): print("  %-22s %8.2f GB %8.3f ms %6.0f GB/s"%(k,b/1e9,t,b/t/1e6 if t else 0))
PY
