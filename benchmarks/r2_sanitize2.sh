set -u
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mixed_stream or golden_fixture or (delete_stream and 12) or (insert_stream and 12) or small_batch_path_vs" > gpurun_out/san_memcheck2.log 2>&1; echo "memcheck exit $?"
grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/san_memcheck2.log | tail -5
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 30 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "(mixed_stream and 20000-20000) or (delete_stream and 12-30000)" > gpurun_out/san_racecheck2.log 2>&1; echo "racecheck exit $?"
grep -E "RACECHECK SUMMARY|hazard|passed|failed" gpurun_out/san_racecheck2.log | sort | uniq -c | tail -12
