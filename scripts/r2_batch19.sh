#!/bin/bash
# Round-2 batch 19: quarter-wise TMEM drain in the tcgen05 epilogue; timeline; correctness subset.
mkdir -p gpurun_out
echo "== pytest gpu (kernels, tc, edge, models)"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tc.py tests/test_gpu_edge.py tests/test_gpu_models.py -m gpu -q --timeout 600 -x 2>&1 | tail -6
timeout 300 python scripts/trace_chain.py 8192 gpurun_out/b19_trace_c2.csv 2>&1 | tail -9
run() { env "$@" timeout 300 python scripts/time_lml.py ${N:-8192} 10 "$*" 2>&1 | tail -1 | tee -a gpurun_out/b19_ab.txt; }
run X=default
run X=default2
N=4096 run X=default
