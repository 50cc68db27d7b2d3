#!/bin/bash
# Round-2 batch 16: device timeline of one C2 evaluation (gpk_debug_trace) + timing.
mkdir -p gpurun_out
timeout 300 python scripts/trace_chain.py 8192 gpurun_out/b16_trace_c2.csv 2>&1 | tail -12
timeout 300 python scripts/time_lml.py 8192 10 default 2>&1 | tail -1
