#!/bin/bash
# final round-2 records (1 GPU): the default bench line, per-op table, ncu launch list, ncu --set full of the big SpMV launches.
# (The full GPU suite, smoke and the 3D 160^3 check ran in the previous call: 531 passed, 1 skipped; its ncu reports were too big
#  to be copied back, hence this leaner repeat.)
cd "$(dirname "$0")/../.." ; mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/rf_bench.json 2> gpurun_out/rf_bench.log; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/rf_bench.json'));print(d['ms_per_step'], d['roofline']['frac'], d['launches_per_cycle'], d['parity']['rel_l2'], d['e2e']['ms_per_step'], d['cpu_baseline']['ms_per_cycle'], d['clocks'])"
B="python bench.py --size 4096 --steps 30 --warmup 5 --no-cpu-baseline --no-parity"
timeout 600 $B --dump-ops gpurun_out/rf_ops_4096.csv > gpurun_out/rf_b2.json 2> gpurun_out/rf_b2.log; echo "ops rc=$?"
N=$(cat gpurun_out/rf_ops_4096.csv.nspmv); echo "spmv launches per cycle: $N"
P="$B --profile-one-cycle"
$P > gpurun_out/rf_plain_profile.log 2>&1 && \
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/rf_launches_4096.csv $P > gpurun_out/rf_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:spmv_ -c 10 -o gpurun_out/rf_prof_down $P > gpurun_out/rf_ncu_down.log 2>&1; echo "ncu down rc=$?"
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:spmv_ -s $((N-10)) -c 10 -o gpurun_out/rf_prof_up $P > gpurun_out/rf_ncu_up.log 2>&1; echo "ncu up rc=$?"
ls -la gpurun_out/rf_*.ncu-rep
for r in down up; do ncu -i gpurun_out/rf_prof_$r.ncu-rep --page raw --csv > gpurun_out/rf_prof_${r}_raw.csv 2>/dev/null; done
SZ=$(du -sm gpurun_out | cut -f1); echo "gpurun_out: $SZ MB"
if [ "$SZ" -gt 55 ]; then rm -f gpurun_out/rf_prof_up.ncu-rep; echo "dropped the up report (raw csv kept)"; fi
SZ=$(du -sm gpurun_out | cut -f1); if [ "$SZ" -gt 55 ]; then rm -f gpurun_out/rf_prof_down.ncu-rep; echo "dropped the down report (raw csv kept)"; fi
du -sm gpurun_out
