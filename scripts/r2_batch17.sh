#!/bin/bash
# Round-2 batch 17: panel kernel with cp.async staging + 8-warp critical update; device timeline; correctness subset.
mkdir -p gpurun_out
echo "== pytest gpu (kernels, tc, edge, models, grad)"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tc.py tests/test_gpu_edge.py tests/test_gpu_models.py tests/test_gpu_grad.py -m gpu -q --timeout 600 -x 2>&1 | tail -4
timeout 300 python scripts/trace_chain.py 8192 gpurun_out/b17_trace_c2.csv 2>&1 | tail -9
run() { env "$@" timeout 300 python scripts/time_lml.py ${N:-8192} 10 "$*" 2>&1 | tail -1 | tee -a gpurun_out/b17_ab.txt; }
run X=default
run X=default2
run GPK_TC_CLUSTER=1
run GPK_TC_CLUSTER=4
N=4096 run X=default
