#!/bin/bash
# 4 GPUs: default exchange (fused push, p2p=2) with the oracle parity check on every rank, and the NCCL exchange for comparison
cd "$(dirname "$0")/../.." ; mkdir -p gpurun_out
nvidia-smi -L | head -4
run() { local name=$1; shift
  timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) bench.py --gpus 4 --size 4096 --steps 30 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r16_$name.json 2> gpurun_out/r16_$name.log; echo "$name rc $?"
}
run n4_default
run n4_p2p0 --opt p2p=0 --no-parity
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r16_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-32s ms %.3f e2e_ms %.3f launches %d parity %s | %s"%(f,d["ms_per_step"],d["e2e"]["ms_per_step"],d["launches_per_cycle"],(d.get("parity") or {}),d["details"]["partition"]))
    except Exception as e:
        print(f,"ERR",e); print(open(f.replace(".json",".log")).read()[-1500:])
PY
