# parity (tests/test_gpu_parity.py) + C2..C5 of the current build; usage: r2_k.sh <tag> [name:ENV=VAL ...]
set -u
tag=${1:-r2k}; shift || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${tag}_pytest.log
if [ $# -eq 0 ]; then set -- cur:PPCSR_X=0; fi
REPS=${REPS:-1} C4=1 C3=1 C5=1 bash benchmarks/run_ab2.sh $tag "$@"
