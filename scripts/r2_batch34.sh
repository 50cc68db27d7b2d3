#!/bin/bash
# Round-2 batch 34: fp32-by-way-of-fp64 factorisation with the scratch inside the caller's workspace: GPU suite, C4 / C3 bench lines.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/b34_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/b34_pytest.log
for w in svgp_c4 sgpr_c3; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 3 --no-svgp > gpurun_out/b34_bench_$w.json 2> gpurun_out/b34_bench_$w.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/b34_bench_$w.json'))
print('$w', {k:d[k] for k in ('value','ms_per_step','objective_vs_cpu_rel_err')}, 'e2e', d['e2e']['value'])
PY
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
