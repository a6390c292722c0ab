#!/bin/bash
# usage: bash tools/gpu/r02_scale.sh "N1 N2 ..."  -- NCCL parity test (2 ranks) + the bench at each N (4096^2, parity checked on every rank)
cd "$(dirname "$0")/../.." ; mkdir -p gpurun_out
NS=$1
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/dist_nccl_check.py > gpurun_out/r12_nccl_check.log 2>&1; echo "nccl check rc=$?"; tail -3 gpurun_out/r12_nccl_check.log
timeout 900 python bench.py --size 4096 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r12_scale_n1.json 2> gpurun_out/r12_scale_n1.log; echo "N=1 rc $?"
for N in $NS; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2962$N bench.py --gpus $N --size 4096 --steps 30 --warmup 5 > gpurun_out/r12_scale_n$N.json 2> gpurun_out/r12_scale_n$N.log; echo "N=$N rc $?"
  grep -h "parity\|PARITY\|Error\|error" gpurun_out/r12_scale_n$N.log | tail -3
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r12_scale_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-36s N=%d ms %.3f DOF/s %.3e e2e_ms %.3f frac %.3f launches %d xg %s parity %s"%(f,d["n_gpus"],d["ms_per_step"],d["value"],d["e2e"]["ms_per_step"],d["roofline"]["frac"],d["launches_per_cycle"],d["details"].get("exchange_groups_per_cycle_rank0"),d["parity"]))
    except Exception as e:
        print(f,"ERR",e); print(open(f.replace(".json",".log")).read()[-1500:])
PY
