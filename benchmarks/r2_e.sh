#!/bin/bash
# round 2: parity tests incl. the small-batch path, the batch-size sweep (C5), C4/C2 lines
tag=${1:-r2e}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/${tag}_pytest.log
timeout 600 python benchmarks/sweep.py --scale 20 --max-batch 10000000 --reps 8 > gpurun_out/${tag}_sweep.jsonl 2> gpurun_out/${tag}_sweep.err; echo "sweep exit $?"
cat gpurun_out/${tag}_sweep.jsonl | cut -c1-400
PPCSR_SPARSE=never timeout 600 python benchmarks/sweep.py --scale 20 --max-batch 100000 --reps 8 > gpurun_out/${tag}_sweep_nosparse.jsonl 2>/dev/null
cat gpurun_out/${tag}_sweep_nosparse.jsonl | cut -c1-300
timeout 600 python bench.py --only-headline --steps 5 > gpurun_out/${tag}_c4.json 2> gpurun_out/${tag}_c4.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_c4.json"))
    print("C4", round(d["value"]/1e9,3), "G upd/s e2e", round(d["e2e"]["value"]/1e9,3), {k:round(v,3) for k,v in d["stages_ms"].items()}, round(d["roofline"]["frac"],4), d["parity"]["golden"]["match"])
except Exception as e: print("failed", e)
PY
