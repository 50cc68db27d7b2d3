#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/leaf_timing.py 2>&1 | tail -8 | tee gpurun_out/b23_leaf.txt
