#!/bin/bash
# round 2: A/B of the locate window capacity (alternate libraries), same box
tag=${1:-r2d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/${tag}_pytest.log
for lib in default cap4096 cap2048; do
  if [ $lib = default ]; then unset PPCSR_B200_LIB; else export PPCSR_B200_LIB=$PWD/parallel-packed-csr_b200/libppcsr_b200_$lib.so; fi
  for cfg in C4 C2 C3; do
    timeout 600 python bench.py --config $cfg --only-headline --steps 5 --e2e-steps 5 > gpurun_out/${tag}_${lib}_$cfg.json 2> gpurun_out/${tag}_${lib}_$cfg.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_${lib}_$cfg.json"))
    print("$lib $cfg", round(d["value"]/1e9,3), "G upd/s e2e", round(d["e2e"]["value"]/1e9,3), round(d["e2e"]["ms_per_step"],2), round(d["e2e"]["ms_per_step_unpipelined"],2), {k:round(v,3) for k,v in d["stages_ms"].items()}, round(d["roofline"]["frac"],4))
except Exception as e: print("$lib $cfg failed", e)
PY
  done
done
unset PPCSR_B200_LIB
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^k_' -c 200 --csv \
    --log-file gpurun_out/${tag}_c4_launches.csv python bench.py --only-headline --steps 1 --warmup 1 --e2e-steps 1 \
    > gpurun_out/${tag}_c4_launches.log 2>&1; echo "ncu launches exit $?"
