#!/bin/bash
# Quick GPU check: parity tests + bench (+ optional extra command in $1)
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['cpu_baseline']['value'])
print('roofline', d['roofline']['achieved'], d['roofline']['pipe_frac_nominal'], 'kbuild', d['kbuild_roofline']['frac'], d['kbuild_roofline']['ms_per_step'])
print(d['kernel_classes'])
PY
tail -3 gpurun_out/bench.err
if [ -n "$1" ]; then echo "== extra: $1"; bash -c "$1"; fi
