#!/bin/bash
# Round-2 batch 25 (final code): full GPU suite and the bench lines of all five BASELINE configs.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/b25_pytest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/b25_pytest.log | cut -c1-200
for w in gpr_c2 gpr_c1 sgpr_c3 svgp_c4 gpr_c5; do
  extra="--no-svgp"; [ $w = gpr_c2 ] && extra=""
  echo "== bench $w"; timeout 900 python bench.py --workload $w --steps 20 --warmup 3 $extra > gpurun_out/b25_bench_$w.json 2> gpurun_out/b25_bench_$w.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/b25_bench_$w.json'))
print('$w', {k:d[k] for k in ('value','ms_per_step','gpu_launches','objective_vs_cpu_rel_err')}, 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], 'frac', d['roofline']['frac'], d.get('value_and_grad'), d.get('posterior_predict'))
PY
done
