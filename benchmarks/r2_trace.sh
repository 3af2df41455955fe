# development: per-warp clock stamps of k_rebalance_m (library built with -DPPCSR_M_TRACE); usage: r2_trace.sh <tag> <configs...>
set -u
tag=$1; shift
for cfg in "$@"; do
  PPCSR_B200_LIB=$PWD/parallel-packed-csr_b200/libppcsr_b200_trace.so PPCSR_TRACE_OUT=$PWD/gpurun_out/${tag}_$cfg.bin \
    python bench.py --config $cfg --only-headline --no-cpu-baseline --steps 1 --warmup 1 --e2e-steps 1 > gpurun_out/${tag}_$cfg.json 2> gpurun_out/${tag}_$cfg.err
  echo "$cfg exit $?"; ls -la gpurun_out/${tag}_$cfg.bin
done
