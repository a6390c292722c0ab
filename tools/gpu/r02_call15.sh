#!/bin/bash
# 2 GPUs: where the N=2 cycle goes with the fused push (per-op table of rank 0, agglomeration threshold, PDL)
cd "$(dirname "$0")/../.." ; mkdir -p gpurun_out
run() { # name, extra args
  local name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) bench.py --gpus 2 --size 4096 --steps 30 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r15_$name.json 2> gpurun_out/r15_$name.log; echo "$name rc $?"
}
run p2p2 --opt p2p=2 --dump-ops gpurun_out/r15_ops_n2_p2p2.csv
run p2p2_agg64k --opt p2p=2 --opt agg_rows=65536 --no-parity
run p2p2_agg1m --opt p2p=2 --opt agg_rows=1048576 --no-parity
run p2p2_nopdl --opt p2p=2 --opt pdl=0 --no-parity
run p2p2_nograph --opt p2p=2 --opt graph=0 --no-parity
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r15_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-36s ms %.3f e2e_ms %.3f launches %d parity %s"%(f,d["ms_per_step"],d["e2e"]["ms_per_step"],d["launches_per_cycle"],(d.get("parity") or {}).get("rel_l2")))
    except Exception as e:
        print(f,"ERR",e); print(open(f.replace(".json",".log")).read()[-800:])
PY
