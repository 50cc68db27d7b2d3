#!/bin/bash
# Round-2 batch 35: C4 end-to-end A/B of the fp32-by-way-of-fp64 factorisation (host-side cost of the extra launches).
mkdir -p gpurun_out
for v in 1 0 1 0; do
  GPK_F32_VIA_F64=$v timeout 600 python bench.py --workload svgp_c4 --steps 20 --warmup 3 --no-svgp 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('via_f64=$v', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'launches/step', d['gpu_launches']/20)"
done
