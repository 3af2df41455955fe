set -u
mkdir -p gpurun_out
for sc_b in "21 12500000" "22 25000000" "23 50000000"; do set -- $sc_b
  python bench.py --scale $1 --batch $2 --workload skewed --only-headline --no-cpu-baseline --steps 5 --e2e-steps 1 > gpurun_out/r2w_$1.json 2>/dev/null
  python - <<PY
import json
j=json.loads(open("gpurun_out/r2w_$1.json").read().strip().splitlines()[-1]); s=j["stages_ms"]; b=j["batch"]
print("scale $1 batch $2: %.2f G/s total %.3f sort %.3f locate %.3f select %.3f reb %.3f (kernel %.3f) slots %d->%d whole %s launches %s"%(j["value"]/1e9,s["ms_total"],s["ms_sort"],s["ms_locate"],s["ms_select"],s["ms_rebalance"],s["ms_rebalance_kernel"],b["slots_before"],b["slots_after"],b["whole_array"],j.get("gpu_launches")))
PY
done
