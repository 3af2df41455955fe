# round 2, session 2: the rank-dense rebalance kernel (k_rebalance_m) against k_rebalance_p, parity first
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2j_pytest.log
REPS=1 C4=1 C3=1 C5=1 bash benchmarks/run_ab2.sh r2j m:PPCSR_REB_KERNEL=10 p:PPCSR_REB_KERNEL=9
