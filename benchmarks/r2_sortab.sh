set -u
mkdir -p gpurun_out
L=$PWD/parallel-packed-csr_b200
for rep in 1 2; do for v in "$@"; do
  lib=$L/libppcsr_b200.so; [ $v != cur ] && lib=$L/libppcsr_b200_$v.so
  for cfg in "--config C4" "--config C2" "--scale 21 --batch 12500000 --workload skewed"; do
    PPCSR_B200_LIB=$lib python bench.py $cfg --only-headline --no-cpu-baseline --steps 4 --e2e-steps 1 > gpurun_out/sortab.json 2>/dev/null
    python - <<PY
import json
j=json.loads(open("gpurun_out/sortab.json").read().strip().splitlines()[-1]); s=j["stages_ms"]
print("$v [$cfg] G/s %.2f sort %.3f locate %.3f reb %.3f"%(j["value"]/1e9,s["ms_sort"],s["ms_locate"],s["ms_rebalance"]))
PY
  done; done; done
