set -u
mkdir -p gpurun_out
for cfg_cl in "C3 56" "C3 58" "C3 60" "C3 62" "C3 63" "C4 60" "C4 61" "C4 62" "C5 60" "C5 61" "C5 62"; do set -- $cfg_cl
  PPCSR_REB_CL=$2 python bench.py --config $1 --only-headline --no-cpu-baseline --steps 6 --e2e-steps 1 > gpurun_out/r2x_$1_$2.json 2>/dev/null
  python - <<PY
import json
j=json.loads(open("gpurun_out/r2x_$1_$2.json").read().strip().splitlines()[-1])
print("$1 CL=$2 reb_ms %.4f frac %.3f G/s %.2f"%(j["roofline"]["kernel_ms"],j["roofline"]["frac"],j["value"]/1e9))
PY
done
