#!/bin/bash
# Round-2 batch 6: everything so far -- full GPU suite (new: widen, widen2, grad, edge), C2 timing, bench with the SVGP section.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/b6_pytest.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/b6_pytest.log | cut -c1-220
run() { env "$@" timeout 300 python scripts/time_lml.py 8192 8 "$*" 2>&1 | tail -1 | tee -a gpurun_out/b6_ab.txt; }
run X=default
run GPK_TC_MIN_K=512
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/b6_bench.json 2> gpurun_out/b6_bench.err; echo "rc=$?"; tail -5 gpurun_out/b6_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/b6_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])
print('roofline', {k:d['roofline'][k] for k in ('achieved','peak','frac','fp64_equivalent_tflops','kernel_ms_per_step')})
print('dmma', d['roofline']['dmma_class'])
print('vg', d.get('value_and_grad'))
print('svgp', json.dumps(d.get('svgp_c4'), indent=0)[:1500])
PY
