#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/leaf_timing.py 2>&1 | tail -3
for k in 512 256; do
GPK_TC_MIN_K=$k timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('tc_min_k $k', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernel_classes'].items()})"
done
