#!/bin/bash
# Round-2 batch 18: bulk-reduction epilogue of the tcgen05 update, cp.async leaf load / X_top staging; timeline; correctness.
mkdir -p gpurun_out
echo "== pytest gpu (kernels, tc, edge, models, grad)"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tc.py tests/test_gpu_edge.py tests/test_gpu_models.py tests/test_gpu_grad.py -m gpu -q --timeout 600 -x 2>&1 | tail -12
timeout 300 python scripts/trace_chain.py 8192 gpurun_out/b18_trace_c2.csv 2>&1 | tail -9
run() { env "$@" timeout 300 python scripts/time_lml.py ${N:-8192} 10 "$*" 2>&1 | tail -1 | tee -a gpurun_out/b18_ab.txt; }
run X=default
run X=default2
run GPK_TC_SLICES=7
N=4096 run X=default
N=2048 run X=default
N=16384 run X=default
