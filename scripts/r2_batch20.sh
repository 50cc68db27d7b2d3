#!/bin/bash
# Round-2 batch 20: flag hops (panel kernels poll the leaf's completion counter) A/B; correctness subset; timeline.
mkdir -p gpurun_out
echo "== pytest gpu (kernels, tc, edge, models, grad)"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tc.py tests/test_gpu_edge.py tests/test_gpu_models.py tests/test_gpu_grad.py -m gpu -q --timeout 600 -x 2>&1 | tail -6
run() { env "$@" timeout 300 python scripts/time_lml.py ${N:-8192} 10 "$*" 2>&1 | tail -1 | tee -a gpurun_out/b20_ab.txt; }
run X=default
run GPK_FLAG_HOPS=0
run X=default2
N=4096 run X=default
N=4096 run GPK_FLAG_HOPS=0
N=16384 run X=default
N=1000 run X=default
timeout 300 python scripts/trace_chain.py 8192 gpurun_out/b20_trace_c2.csv 2>&1 | tail -9
