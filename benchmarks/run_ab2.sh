# same-box A/B: alternate the kernels / library builds, C2 (and optionally C4), kernel time only
set -u
mkdir -p gpurun_out
tag=${1:-ab2}; shift
for rep in $(seq 1 ${REPS:-2}); do
for v in "$@"; do   # each v: name:ENV=VAL,ENV=VAL
  name=${v%%:*}; envs=${v#*:}
  ( IFS=,; for e in $envs; do export "$e"; done
    python bench.py --config C2 --only-headline --no-cpu-baseline --steps 10 > gpurun_out/${tag}_${name}_c2_$rep.json 2>/dev/null
    if [ "${C4:-0}" = "1" ]; then python bench.py --only-headline --no-cpu-baseline --config C4 --steps 3 > gpurun_out/${tag}_${name}_c4_$rep.json 2>/dev/null; fi
    if [ "${C5:-0}" = "1" ]; then python bench.py --only-headline --no-cpu-baseline --config C5 --steps 10 > gpurun_out/${tag}_${name}_c5_$rep.json 2>/dev/null; fi
    if [ "${C3:-0}" = "1" ]; then python bench.py --only-headline --no-cpu-baseline --config C3 --steps 10 > gpurun_out/${tag}_${name}_c3_$rep.json 2>/dev/null; fi
  )
  python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_${name}_c?_$rep.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "G/s %.2f"%(j["value"]/1e9), "reb_ms %.4f"%j["roofline"]["kernel_ms"], "frac %.3f"%j["roofline"]["frac"])
    except Exception as e: print(f, "failed", e)
PY
done
done
