# usage: bash tools/sweep_kernels.sh N "cfg1" "cfg2" ...   (cfg = space-separated key=value library options)
N=$1; shift
mkdir -p gpurun_out
for cfg in "$@"; do
  o=""; for kv in $cfg; do o="$o --opt $kv"; done
  tag=$(echo $cfg | tr ' =' '__')
  timeout 900 python bench.py --size $N --steps 20 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_${N}_$tag.csv $o > gpurun_out/s_${N}_$tag.json 2> gpurun_out/s_${N}_$tag.log || echo FAIL $cfg
  python - gpurun_out/s_${N}_$tag.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d['roofline']
    print("%-44s ms %.3f cycleGB/s %.0f spmvGB/s %.0f largest %.0f"%(sys.argv[1], d['ms_per_step'], r['cycle_achieved'], r['achieved'], r['largest_launch']['GBps']), flush=True)
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
