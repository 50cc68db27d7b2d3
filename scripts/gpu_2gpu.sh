#!/bin/bash
# 2-GPU runs only (call with gpurun --gpus 2): replicas of C2 and independent minibatches of C4, one process per GPU
mkdir -p gpurun_out
for w in gpr_c2 svgp_c4; do
  echo "== 2 GPUs $w"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --workload $w --steps 10 > gpurun_out/bench2_$w.json 2> gpurun_out/bench2_$w.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench2_$w.json'))
    print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, 'e2e', d['e2e']['value'])
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/bench2_$w.err').read()[-1500:])
PY
done
