#!/bin/bash
# round 2, first GPU pass: parity tests, the new default bench line (C4), reference-arm mechanics at C2 size
tag=${1:-r2a}
mkdir -p gpurun_out
nproc > gpurun_out/${tag}_host.txt; free -g >> gpurun_out/${tag}_host.txt; nvidia-smi -L >> gpurun_out/${tag}_host.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
tail -c 600 gpurun_out/${tag}_bench.err
timeout 300 python bench.py --impl reference --config C2 --steps 1 > gpurun_out/${tag}_ref_c2.json 2> gpurun_out/${tag}_ref_c2.err; echo "ref exit $?"
cat gpurun_out/${tag}_ref_c2.json | cut -c1-600
