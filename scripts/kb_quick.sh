#!/bin/bash
# K-build experiment: parity tests for the builder, then the bench's K-build lines for both register budgets
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 300 -k "kbuild or fast or kdiag or Separate" 2>&1 | tail -5
for mb in 2 3; do
  GPK_KF_MINB=$mb timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kbuild_roofline']
print('minb $mb', 'step', round(d['ms_per_step'],3), 'lower ms', round(k['ms_per_step'],4), 'frac', round(k['frac'],3), 'full ms', round(k['full_matrix']['ms'],4), 'frac', round(k['full_matrix']['frac'],3))"
done
