#!/bin/bash
# Round-2 batch 14: base-256 digit planes (S = 6 + square term / S = 7): full GPU suite, A/B against S = 7 / 8, default bench line.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/b14_pytest.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/b14_pytest.log
run() { env "$@" timeout 300 python scripts/time_lml.py ${N:-8192} 10 "$*" 2>&1 | tail -1 | tee -a gpurun_out/b14_ab.txt; }
run X=default
run GPK_TC_SLICES=6
run GPK_TC_SLICES=7
run GPK_TC_SLICES=8
run GPK_TC_SLICES=6 GPK_TC_CAT=0
run GPK_TC_SLICES=6 GPK_TC_A_TMEM=1
N=4096 run X=default
N=2048 run X=default
N=16384 run X=default
echo "== bench C2"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/b14_bench_c2.json 2> gpurun_out/b14_bench_c2.err; echo "rc=$?"; tail -2 gpurun_out/b14_bench_c2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/b14_bench_c2.json'))
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['slices'], d['roofline']['kernel_ms_per_step'])
PY
