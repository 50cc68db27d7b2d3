#!/bin/bash
# Round-2 batch 24 (rerun for the all-warp pair 0): DMMA block-pair updates / level-32 inverse products inside the leaf; leaf phases; GPU suite; timings.
mkdir -p gpurun_out
timeout 120 python scripts/leaf_timing.py 2>&1 | tail -5 | tee gpurun_out/b24_leaf.txt
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/b24_pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/b24_pytest.log
run() { env "$@" timeout 300 python scripts/time_lml.py ${N:-8192} 10 "$*" 2>&1 | tail -1 | tee -a gpurun_out/b24_ab.txt; }
run X=default
run X=default2
N=4096 run X=default
timeout 300 python scripts/trace_chain.py 8192 gpurun_out/b24_trace_c2.csv 2>&1 | tail -3
