#!/bin/bash
# Round-2 batch 29: one-shuffle 32x32 factorisation (variant 3), programmatic dependent launch of the update: leaf phases, GPU suite, A/B.
mkdir -p gpurun_out
timeout 120 python scripts/leaf_timing.py 2>&1 | tail -3 | tee gpurun_out/b29_leaf.txt
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/b29_pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/b29_pytest.log
run() { env "$@" timeout 300 python scripts/time_lml.py ${N:-8192} 10 "$*" 2>&1 | tail -1 | tee -a gpurun_out/b29_ab.txt; }
run X=default
run GPK_TC_PDL=0
run X=default2
N=4096 run X=default
N=12288 run X=default
timeout 300 python scripts/trace_chain.py 8192 gpurun_out/b29_trace_c2.csv 2>&1 | grep "panel_last\|syrk_start\|leaves"
