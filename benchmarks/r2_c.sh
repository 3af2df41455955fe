#!/bin/bash
# round 2: parity tests + default bench line (C4 with other_configs)
tag=${1:-r2c}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
tail -c 1500 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_bench.json"))
    print("C4", round(d["value"]/1e9,3), "G upd/s e2e", round(d["e2e"]["value"]/1e9,3), {k:round(v,3) for k,v in d["stages_ms"].items()}, round(d["roofline"]["frac"],4), d["parity"]["golden"])
    for k,v in d.get("other_configs",{}).items():
        print(k, round(v["value"]/1e9,3), {a:round(b,3) for a,b in v["stages_ms"].items()}, round(v["roofline_frac"],4)) if "value" in v else print(k, v)
except Exception as e: print("failed", e)
PY
