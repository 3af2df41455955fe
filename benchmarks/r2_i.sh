#!/bin/bash
# round 2: A/B of the rebalance kernel geometry: 256 threads x 2048-slot chunks x 4 CTAs/SM (default) against
# 128 threads x 1024-slot chunks x 7 CTAs/SM (libppcsr_b200_t128.so), same box
tag=${1:-r2i}
mkdir -p gpurun_out
export PPCSR_B200_LIB=$PWD/parallel-packed-csr_b200/libppcsr_b200_t128.so
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or insert_stream or delete_stream or mixed_stream or hub" > gpurun_out/${tag}_pytest_t128.log 2>&1; echo "pytest t128 exit $?"
tail -3 gpurun_out/${tag}_pytest_t128.log
for rep in 1 2; do
for lib in default t128; do
  if [ $lib = default ]; then unset PPCSR_B200_LIB; else export PPCSR_B200_LIB=$PWD/parallel-packed-csr_b200/libppcsr_b200_$lib.so; fi
  for cfg in C4 C2 C3 C5; do
    timeout 600 python bench.py --config $cfg --only-headline --steps 8 --e2e-steps 2 > gpurun_out/${tag}_${lib}_${cfg}_$rep.json 2> gpurun_out/${tag}_${lib}_${cfg}_$rep.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_${lib}_${cfg}_$rep.json"))
    print("$lib $cfg", round(d["value"]/1e9,3), "G upd/s", "reb kernel ms", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],4), d["parity"]["golden"] and d["parity"]["golden"]["match"])
except Exception as e: print("$lib $cfg failed", e)
PY
  done
done
done
