#!/bin/bash
# Round-2 batch 28 (2 GPUs, final code): NCCL multi-rank parity test + the default bench line at N = 2 (C2 replicas + SVGP modes).
mkdir -p gpurun_out
echo "== multirank test"; timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q --timeout 500 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/b28_bench2.json 2> gpurun_out/b28_bench2.err; echo "rc=$?"; tail -3 gpurun_out/b28_bench2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/b28_bench2.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'])
for k,v in d['svgp_c4'].items():
    if isinstance(v,dict): print(k, round(v['evals_per_s'],1), round(v['ms_per_step'],3), v.get('sum_of_shares_vs_full_rel_err'))
PY
