#!/bin/bash
# final round-2 call (1 GPU): full GPU test suite, smoke, the default bench line, per-op table, ncu launch list, ncu --set full of
# the big SpMV launches (-> profiles/r02_ncu_traffic.json via tools/make_traffic_json.py), 3D lAIR check of the long-row kernel
cd "$(dirname "$0")/../.." ; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/rf_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/rf_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/rf_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/rf_smoke.log
timeout 900 python bench.py > gpurun_out/rf_bench.json 2> gpurun_out/rf_bench.log; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/rf_bench.json'));print(d['ms_per_step'], d['roofline']['frac'], d['launches_per_cycle'], d['parity']['rel_l2'], d['e2e']['ms_per_step'], d['cpu_baseline'], d['clocks'])"
B="python bench.py --size 4096 --steps 30 --warmup 5 --no-cpu-baseline --no-parity"
timeout 600 $B --dump-ops gpurun_out/rf_ops_4096.csv > gpurun_out/rf_b2.json 2> gpurun_out/rf_b2.log; echo "ops rc=$?"
N=$(cat gpurun_out/rf_ops_4096.csv.nspmv); echo "spmv launches per cycle: $N"
P="$B --profile-one-cycle"
$P > gpurun_out/rf_plain_profile.log 2>&1 && \
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/rf_launches_4096.csv $P > gpurun_out/rf_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:spmv_ -c 12 -o gpurun_out/rf_prof_down $P > gpurun_out/rf_ncu_down.log 2>&1; echo "ncu down rc=$?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:spmv_ -s $((N-14)) -c 14 -o gpurun_out/rf_prof_up $P > gpurun_out/rf_ncu_up.log 2>&1; echo "ncu up rc=$?"
ls -la gpurun_out/rf_*.ncu-rep
timeout 900 python bench.py --workload adv_diff_fd_3d_lair --size 160 --steps 20 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/rf_ops_3d160.csv > gpurun_out/rf_b3d160.json 2> gpurun_out/rf_b3d160.log; echo "3d rc=$?"
python -c "import json;d=json.load(open('gpurun_out/rf_b3d160.json'));print('3D 160^3 lAIR', d['ms_per_step'], d['roofline']['frac'], d['launches_per_cycle'], d['parity']['rel_l2'])"
