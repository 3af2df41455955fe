#!/bin/bash
# round 2: C4 strong scaling on N GPUs (usage: r2_mg.sh <tag> <N> [extra bench args])
tag=$1; N=$2; shift; shift
mkdir -p gpurun_out
nvidia-smi -L | head -8
PPCSR_ROUTE_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/${tag}_n$N.json 2> gpurun_out/${tag}_n$N.err; echo "bench exit $?"
grep -E "^\[rank|Error|error|Traceback" gpurun_out/${tag}_n$N.err | cut -c1-400 | head -40
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_n$N.json"))
    print("N=$N", round(d["value"]/1e9,3), "G upd/s", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"]/1e9,3), {k:round(v,3) for k,v in d["stages_ms"].items()}, d["parity"]["golden"] and d["parity"]["golden"]["match"], "pagerank ms", d.get("pagerank",{}).get("ms_per_step"))
except Exception as e: print("failed", e)
PY
