#!/bin/bash
tag=${1:-r2g}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?"
tail -6 gpurun_out/${tag}_pytest.log
for cfg in C4 C2 C3 C5; do
timeout 600 python bench.py --config $cfg --only-headline --steps 5 > gpurun_out/${tag}_$cfg.json 2> gpurun_out/${tag}_$cfg.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_$cfg.json"))
    print("$cfg", round(d["value"]/1e9,3), "G upd/s e2e", round(d["e2e"]["value"]/1e9,3), {k:round(v,3) for k,v in d["stages_ms"].items()}, round(d["roofline"]["frac"],4), d["parity"]["golden"] and d["parity"]["golden"]["match"])
except Exception as e: print("failed", e)
PY
done
timeout 600 python benchmarks/sweep.py --scale 20 --max-batch 100000 --reps 8 --workload insert --no-pagerank > gpurun_out/${tag}_sweep_ins.jsonl 2>/dev/null
cat gpurun_out/${tag}_sweep_ins.jsonl | cut -c1-200
