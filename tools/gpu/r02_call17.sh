#!/bin/bash
# 2 GPUs: 3D lAIR 192^3 (7.1 M rows) at N = 1 and N = 2 -- the bandwidth-bound regime of the strong-scaling curve
cd "$(dirname "$0")/../.." ; mkdir -p gpurun_out
W="--workload adv_diff_fd_3d_lair --size 192 --steps 20 --warmup 3 --no-cpu-baseline"
timeout 900 python bench.py $W > gpurun_out/r17_3d192_n1.json 2> gpurun_out/r17_3d192_n1.log; echo "N=1 rc $?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29917 bench.py --gpus 2 $W > gpurun_out/r17_3d192_n2.json 2> gpurun_out/r17_3d192_n2.log; echo "N=2 rc $?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r17_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-32s N=%d ms %.3f e2e_ms %.3f frac %.3f launches %d parity %s"%(f,d["n_gpus"],d["ms_per_step"],d["e2e"]["ms_per_step"],d["roofline"]["frac"],d["launches_per_cycle"],(d.get("parity") or {}).get("rel_l2")))
    except Exception as e:
        print(f,"ERR",e); print(open(f.replace(".json",".log")).read()[-1500:])
PY
grep "\[bench\]" gpurun_out/r17_3d192_n1.log | tail -4
