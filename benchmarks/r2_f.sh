#!/bin/bash
# round 2: parity tests, batch-size sweep with the one-launch small sort, C4 / C2 lines
tag=${1:-r2f}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?"
tail -12 gpurun_out/${tag}_pytest.log
timeout 600 python benchmarks/sweep.py --scale 20 --max-batch 1000000 --reps 8 > gpurun_out/${tag}_sweep.jsonl 2> gpurun_out/${tag}_sweep.err; echo "sweep exit $?"
cat gpurun_out/${tag}_sweep.jsonl | cut -c1-420
timeout 600 python benchmarks/sweep.py --scale 20 --max-batch 100000 --reps 8 --workload insert --no-pagerank > gpurun_out/${tag}_sweep_ins.jsonl 2>/dev/null
cat gpurun_out/${tag}_sweep_ins.jsonl | cut -c1-300
for cfg in C4 C2; do
timeout 600 python bench.py --config $cfg --only-headline --steps 5 > gpurun_out/${tag}_$cfg.json 2> gpurun_out/${tag}_$cfg.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_$cfg.json"))
    print("$cfg", round(d["value"]/1e9,3), "G upd/s e2e", round(d["e2e"]["value"]/1e9,3), {k:round(v,3) for k,v in d["stages_ms"].items()}, round(d["roofline"]["frac"],4), d["parity"]["golden"] and d["parity"]["golden"]["match"])
except Exception as e: print("failed", e)
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^k_' -c 400 --csv \
    --log-file gpurun_out/${tag}_small_launches.csv python benchmarks/sweep.py --scale 20 --max-batch 10000 --reps 2 --no-pagerank \
    > gpurun_out/${tag}_small_launches.log 2>&1; echo "ncu exit $?"
