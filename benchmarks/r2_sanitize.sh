# compute-sanitizer over a few small parity cases (memcheck + racecheck of the shared-memory protocol)
set -u
mkdir -p gpurun_out
T="tests/test_gpu_parity.py::test_long_insert_runs_one_value tests/test_gpu_parity.py::test_hub_vertex_grow_and_shrink"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest $T -m gpu -x -q > gpurun_out/san_memcheck.log 2>&1; echo "memcheck exit $?"
grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/san_memcheck.log | tail -5
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 30 python -m pytest "tests/test_gpu_parity.py::test_long_insert_runs_one_value" -m gpu -x -q > gpurun_out/san_racecheck.log 2>&1; echo "racecheck exit $?"
grep -E "RACECHECK SUMMARY|hazard|passed|failed" gpurun_out/san_racecheck.log | sort | uniq -c | tail -12
