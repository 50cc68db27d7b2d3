#!/bin/bash
# Round-2 batch 33: fp32 factorisations by way of the fp64 path (GPK_F32_VIA_F64): GPU suite, C3 / C4 bench lines, A/B.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/b33_pytest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/b33_pytest.log
for w in sgpr_c3 svgp_c4; do
  echo "== bench $w"; timeout 900 python bench.py --workload $w --steps 20 --warmup 3 --no-svgp > gpurun_out/b33_bench_$w.json 2> gpurun_out/b33_bench_$w.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/b33_bench_$w.json'))
print('$w', {k:d[k] for k in ('value','ms_per_step','objective_vs_cpu_rel_err')}, 'e2e', d['e2e']['value'], {k:round(v['ms_per_step'],3) for k,v in d['kernel_classes'].items()})
PY
done
GPK_F32_VIA_F64=0 timeout 600 python bench.py --workload svgp_c4 --steps 20 --warmup 3 --no-svgp 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('svgp_c4 fp32 kernels', d['ms_per_step'])"
