# round-end measurement set on one GPU: parity tests, default bench line (+ reference arm), C3/C4/C5 lines, ncu launch list
set -u
tag=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?"; tail -n 1 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference exit $?"
timeout 300 python bench.py --no-cpu-baseline --config C3 > gpurun_out/${tag}_c3_delete.json 2>/dev/null
timeout 300 python bench.py --no-cpu-baseline --config C5 > gpurun_out/${tag}_c5_mixed.json 2>/dev/null
timeout 400 python bench.py --no-cpu-baseline --config C4 --steps 3 > gpurun_out/${tag}_c4_skewed24.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^k_' -c 800 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_launches.log 2>&1; echo "ncu exit $?"
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j.get("roofline",{}); print(f, "G/s %.3f"%(j["value"]/1e9), "e2e %.3f"%(j["e2e"]["value"]/1e9), "reb_ms", r.get("kernel_ms"), "frac", r.get("frac"), j.get("cpu_baseline"))
    except Exception as e: print(f, "failed", e)
PY
