set -u
mkdir -p gpurun_out
[ "${SKIP_TESTS:-0}" = 1 ] || timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2s_pytest.log
L=$PWD/parallel-packed-csr_b200
for rep in 1 2; do for v in "$@"; do
  lib=$L/libppcsr_b200.so; [ $v != cur ] && lib=$L/libppcsr_b200_$v.so
  for cfg in C4 C2 C3; do
    PPCSR_B200_LIB=$lib python bench.py --config $cfg --only-headline --no-cpu-baseline --steps 4 --e2e-steps 1 > gpurun_out/r2s_${v}_${cfg}_$rep.json 2>/dev/null
    python - <<PY
import json
j=json.loads(open("gpurun_out/r2s_${v}_${cfg}_$rep.json").read().strip().splitlines()[-1]); s=j["stages_ms"]
print("$v $cfg G/s %.2f sort %.3f locate %.3f select %.3f reb %.3f"%(j["value"]/1e9,s["ms_sort"],s["ms_locate"],s["ms_select"],s["ms_rebalance"]))
PY
  done; done; done
