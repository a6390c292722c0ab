#!/bin/bash
# 2 GPUs: fused peer-memory push (p2p=2) -- cluster tests on one GPU, two-process parity, bench at N=2 for p2p=0/1/2
cd "$(dirname "$0")/../.." ; mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/r14_dist.log 2>&1; echo "dist tests rc=$?"; tail -3 gpurun_out/r14_dist.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "long" > gpurun_out/r14_long.log 2>&1; echo "long-row tests rc=$?"; tail -2 gpurun_out/r14_long.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tests/dist_nccl_check.py > gpurun_out/r14_nccl_check.log 2>&1; echo "nccl check rc=$?"; tail -3 gpurun_out/r14_nccl_check.log
for P in 2 0 1; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2972$P bench.py --gpus 2 --size 4096 --steps 30 --warmup 5 --no-cpu-baseline --opt p2p=$P > gpurun_out/r14_n2_p2p$P.json 2> gpurun_out/r14_n2_p2p$P.log; echo "N=2 p2p=$P rc $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r14_n2_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-32s ms %.3f e2e_ms %.3f launches %d xg %s parity %s"%(f,d["ms_per_step"],d["e2e"]["ms_per_step"],d["launches_per_cycle"],d["details"].get("exchange_groups_per_cycle_rank0"),d["parity"]["rel_l2"]))
    except Exception as e:
        print(f,"ERR",e); print(open(f.replace(".json",".log")).read()[-1500:])
PY
