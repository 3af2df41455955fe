# A/B of the rebalance kernels on the GPU box: parity tests, then C2 / C3 / C4 bench lines per kernel
set -u
mkdir -p gpurun_out
tag=${1:-ab}; kernels=${2:-"6"}; full=${3:-0}
for k in $kernels; do
  PPCSR_REB_KERNEL=$k timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_k$k.log 2>&1; echo "pytest k$k exit $?"; tail -2 gpurun_out/${tag}_pytest_k$k.log
  PPCSR_REB_KERNEL=$k timeout 300 python bench.py --config C2 --no-cpu-baseline > gpurun_out/${tag}_c2_k$k.json 2> gpurun_out/${tag}_c2_k$k.err; echo "bench k$k exit $?"
  python - <<PY
import json
for f in ["gpurun_out/${tag}_c2_k$k.json"]:
    j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, j["value"]/1e9, j["roofline"]["kernel_ms"], j["roofline"]["frac"], j["stages_ms"])
PY
  if [ "$full" = "1" ]; then
    PPCSR_REB_KERNEL=$k timeout 300 python bench.py --no-cpu-baseline --config C3 > gpurun_out/${tag}_c3_k$k.json 2>/dev/null
    PPCSR_REB_KERNEL=$k timeout 300 python bench.py --no-cpu-baseline --config C5 > gpurun_out/${tag}_c5_k$k.json 2>/dev/null
    PPCSR_REB_KERNEL=$k timeout 400 python bench.py --no-cpu-baseline --config C4 --steps 3 > gpurun_out/${tag}_c4_k$k.json 2>/dev/null
    python - <<PY
import json
for f in ["gpurun_out/${tag}_c3_k$k.json","gpurun_out/${tag}_c5_k$k.json","gpurun_out/${tag}_c4_k$k.json"]:
    j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, j["value"]/1e9, j["roofline"]["kernel_ms"], j["roofline"]["frac"], j["stages_ms"])
PY
  fi
done
