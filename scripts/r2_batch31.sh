#!/bin/bash
# Round-2 batch 31: division-free / line-coalesced hi-lo split pre-pass of the tf32 GEMM: GPU suite, C3 / C4 bench lines.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/b31_pytest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/b31_pytest.log
for w in sgpr_c3 svgp_c4; do
  echo "== bench $w"; timeout 900 python bench.py --workload $w --steps 20 --warmup 3 --no-svgp > gpurun_out/b31_bench_$w.json 2> gpurun_out/b31_bench_$w.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/b31_bench_$w.json'))
print('$w', {k:d[k] for k in ('value','ms_per_step','objective_vs_cpu_rel_err')}, 'e2e', d['e2e']['value'], {k:round(v['ms_per_step'],3) for k,v in d['kernel_classes'].items()}, d.get('posterior_predict'))
PY
done
