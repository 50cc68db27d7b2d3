#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_models.py tests/test_gpu_kernels.py -m gpu -q -x --timeout 300 2>&1 | tail -3
for c in 2 1; do
for w in svgp_c4 sgpr_c3; do
GPK_TF32_CLUSTER=$c timeout 300 python bench.py --workload $w --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cluster $c $w', round(d['ms_per_step'],3), 'relerr', d['config']['objective_vs_cpu_rel_err'], {k:round(v['ms_per_step'],3) for k,v in d['kernel_classes'].items()})"
done
done
