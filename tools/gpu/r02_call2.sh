#!/bin/bash
# 4096^2: warp-tile kernel bench + option sweep + ncu (launch list and --set full on the big launches)
cd "$(dirname "$0")/../.." ; mkdir -p gpurun_out
B="python bench.py --size 4096 --steps 30 --warmup 5 --no-cpu-baseline"
timeout 1200 $B --dump-ops gpurun_out/r2_ops_4096_k2.csv \
   --compare-opt wt_stages=2 --compare-opt wt_stages=4 --compare-opt wt_stages=3,ctas_per_sm=1 --compare-opt ctas_per_sm=0,epi_classes=0 --compare-opt epi_classes=1,pdl=0 \
   > gpurun_out/r2_b4096_k2.json 2> gpurun_out/r2_b4096_k2.log; echo "bench 4096 k2 rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r2_b4096_k2.json'));print(d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['launches_per_cycle'], d['parity'], d['compare_opt_ms'], d['e2e']['ms_per_step'])"
tail -4 gpurun_out/r2_b4096_k2.log
N=$(cat gpurun_out/r2_ops_4096_k2.csv.nspmv)
P="$B --no-parity --profile-one-cycle"
$P > gpurun_out/r2_plain_profile.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_4096.csv $P > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
$P > gpurun_out/r2_plain_profile.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:spmv_wt -c 8 -o gpurun_out/r2_prof_down $P > gpurun_out/r2_ncu_down.log 2>&1; echo "ncu down rc=$?"
$P > gpurun_out/r2_plain_profile.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:spmv_wt -s $((N-12)) -c 12 -o gpurun_out/r2_prof_up $P > gpurun_out/r2_ncu_up.log 2>&1; echo "ncu up rc=$?"
ls -la gpurun_out/*.ncu-rep
