set -u
mkdir -p gpurun_out
L=$PWD/parallel-packed-csr_b200
run() { # name env lib
  for cfg in C4 C2; do
    env $2 PPCSR_B200_LIB=$3 python bench.py --config $cfg --no-cpu-baseline --steps 2 --warmup 1 --e2e-steps 1 > gpurun_out/r2v_$1_$cfg.json 2>gpurun_out/r2v_$1_$cfg.err
    python - <<PY
import json
j=json.loads(open("gpurun_out/r2v_$1_$cfg.json").read().strip().splitlines()[-1]); p=j.get("pagerank") or {}
print("$1 $cfg G/s %.2f pagerank ms %.3f GB/s %.0f"%(j["value"]/1e9,p.get("ms_per_step",-1),p.get("achieved_gbs",-1)))
PY
  done
}
run prev X=1 $L/libppcsr_b200_prev.so
run ldcs X=1 $L/libppcsr_b200.so
run l2win PPCSR_PR_L2WIN=1 $L/libppcsr_b200.so
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pagerank or oracle" 2>&1 | tail -2
