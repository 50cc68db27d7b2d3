#!/bin/bash
# Round-2 batch 30: what the driver runs at round end -- smoke(), the default bench line, the reference arm.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/b30_bench.json 2> gpurun_out/b30_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/b30_bench.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/b30_ref.json 2> gpurun_out/b30_ref.err; echo "ref rc=$?"; tail -2 gpurun_out/b30_ref.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/b30_bench.json'))
print({k:d[k] for k in ('metric','value','unit','ms_per_step','gpu_launches','vs_baseline','dtype')}, 'e2e', d['e2e'], 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'clocks', d['clocks'])
print('cpu', d['cpu_baseline'])
r=json.load(open('gpurun_out/b30_ref.json'))
print({k:r.get(k) for k in ('impl','metric','value','unit','ms_per_step','config')}, r.get('cpu_baseline'))
PY
