#!/bin/bash
# Round-2 batch 13 (8 GPUs): the default bench line at N = 8 -- C2 replicas + BASELINE configs[3] (SVGP) in every sharding mode.
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/b13_bench8.json 2> gpurun_out/b13_bench8.err; echo "rc=$?"; tail -4 gpurun_out/b13_bench8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/b13_bench8.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'])
print(d['step_stats'])
for k,v in d['svgp_c4'].items():
    if isinstance(v,dict): print(k, round(v['evals_per_s'],1), round(v['ms_per_step'],3), v.get('sum_of_shares_vs_full_rel_err'), [round(x,3) for x in v['step_stats']['per_rank_median_ms']])
PY
