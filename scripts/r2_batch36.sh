#!/bin/bash
# Round-2 batch 36: the default bench line (C2 + the four SVGP sharding modes) with the final code.
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/b36_bench.json 2> gpurun_out/b36_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/b36_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/b36_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'])
for k,v in d['svgp_c4'].items():
    if isinstance(v,dict) and 'evals_per_s' in v: print(k, round(v['evals_per_s'],1), round(v['ms_per_step'],3), v.get('sum_of_shares_vs_full_rel_err'))
PY
