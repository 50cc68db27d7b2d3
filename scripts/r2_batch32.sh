#!/bin/bash
# Round-2 batch 32: flakiness / stress -- the GPU suite three times, 300 back-to-back C2 evaluations, concurrent-stream test x 15.
mkdir -p gpurun_out
for i in 1 2 3; do timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -p no:cacheprovider 2>&1 | tail -1; done
timeout 600 python scripts/time_lml.py 8192 300 stress300 2>&1 | tail -1
for i in $(seq 1 15); do timeout 300 python -m pytest tests/test_gpu_edge.py -m gpu -q -k "concurrent" 2>&1 | tail -1; done | sort | uniq -c
timeout 300 python scripts/time_lml.py 5000 50 n5000 2>&1 | tail -1
timeout 300 python scripts/time_lml.py 8321 50 n8321 2>&1 | tail -1
