#!/bin/bash
# Round-2 batch 8: full GPU suite, bench lines of all five BASELINE configs, launch list and ncu --set full captures.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/b8_pytest.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/b8_pytest.log | cut -c1-200
for w in gpr_c2 gpr_c1 sgpr_c3 svgp_c4 gpr_c5; do
  extra="--no-svgp"; [ $w = gpr_c2 ] && extra=""
  echo "== bench $w"; timeout 900 python bench.py --workload $w --steps 20 --warmup 3 $extra > gpurun_out/b8_bench_$w.json 2> gpurun_out/b8_bench_$w.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/b8_bench_$w.json'))
print('$w', {k:d[k] for k in ('value','ms_per_step','gpu_launches','objective_vs_cpu_rel_err')}, 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], 'frac', d['roofline']['frac'], d.get('value_and_grad'), d.get('posterior_predict'))
PY
done
echo "== ncu launch list (C2)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 210 --csv --log-file gpurun_out/b8_launches_c2.csv python scripts/time_lml.py 8192 1 ncu > gpurun_out/b8_ncu0.log 2>&1; echo "rc=$?"
echo "== ncu --set full: syrk_i8 (second captured evaluation: launches 17..32 -> K=256,256,512,...)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:syrk_i8 -s 31 -c 31 -o gpurun_out/b8_syrk python scripts/time_lml.py 8192 1 ncu > gpurun_out/b8_ncu1.log 2>&1; echo "rc=$?"
echo "== ncu --set full: kbuild_fast, panel, leaf"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"kbuild_fast|potrf_panel|potrf_leaf" -s 130 -c 6 -o gpurun_out/b8_misc python scripts/time_lml.py 8192 1 ncu > gpurun_out/b8_ncu2.log 2>&1; echo "rc=$?"
echo "== ncu --set full: batched tf32 (SVGP)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 40 -c 4 -o gpurun_out/b8_tf32 python bench.py --workload svgp_c4 --steps 3 --no-svgp > gpurun_out/b8_ncu3.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
