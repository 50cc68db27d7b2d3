#!/bin/bash
# Round-2 batch 26: concurrent factorisations (single-poller rule): edge tests, C5 bench, C2 timing.
mkdir -p gpurun_out
echo "== pytest edge + models"; timeout 900 python -m pytest tests/test_gpu_edge.py tests/test_gpu_models.py -m gpu -q --timeout 600 -x 2>&1 | tail -5
w=gpr_c5
echo "== bench $w"; timeout 900 python bench.py --workload $w --steps 20 --warmup 3 --no-svgp > gpurun_out/b25_bench_$w.json 2> gpurun_out/b25_bench_$w.err; echo "rc=$?"; tail -3 gpurun_out/b25_bench_$w.err
python - <<PY
import json
d=json.load(open('gpurun_out/b25_bench_$w.json'))
print('$w', {k:d[k] for k in ('value','ms_per_step','gpu_launches','objective_vs_cpu_rel_err')}, 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], 'frac', d['roofline']['frac'])
PY
timeout 300 python scripts/time_lml.py 8192 10 default 2>&1 | tail -1
