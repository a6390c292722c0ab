# usage: bash tools/scale_run.sh SIZE "N1 N2 ..."   -- strong-scaling runs back to back on one box (hierarchy cached after the first)
SIZE=$1; NS=$2
mkdir -p gpurun_out
python bench.py --size $SIZE --steps 30 --warmup 5 --dump-ops gpurun_out/final_ops_$SIZE.csv > gpurun_out/scale_${SIZE}_n1.json 2> gpurun_out/scale_${SIZE}_n1.log; echo "N=1 rc $?"
for N in $NS; do
  for P2P in 1 0; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --size $SIZE --steps 30 --warmup 5 --opt p2p=$P2P > gpurun_out/scale_${SIZE}_n${N}_p2p$P2P.json 2> gpurun_out/scale_${SIZE}_n${N}_p2p$P2P.log; echo "N=$N p2p=$P2P rc $?"
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/scale_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-44s N=%d ms %.3f DOF/s %.3e e2e_ms %.3f cycleGB/s(rank0) %.0f launches %d xg %s cpu %s"%(f,d["n_gpus"],d["ms_per_step"],d["value"],d["e2e"]["ms_per_step"],d["roofline"]["cycle_achieved"],d["launches_per_cycle"],d["config"].get("exchange_groups_per_cycle_rank0"),(d.get("cpu_baseline") or {}).get("ms_per_cycle")))
    except Exception as e:
        print(f,"ERR",e); print(open(f.replace(".json",".log")).read()[-1500:])
PY
