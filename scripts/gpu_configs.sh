#!/bin/bash
# other BASELINE configs + 2-GPU run (call with gpurun --gpus 2)
mkdir -p gpurun_out
for w in sgpr_c3 svgp_c4 gpr_c5 gpr_c1; do
  echo "== $w"; timeout 600 python bench.py --workload $w --steps 10 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$w.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], 'relerr', d['config']['objective_vs_cpu_rel_err'])
    print(d['kernel_classes'])
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/bench_$w.err').read()[-1500:])
PY
done
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  for w in gpr_c2 svgp_c4; do
  echo "== 2 GPUs $w"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $w --steps 10 > gpurun_out/bench2_$w.json 2> gpurun_out/bench2_$w.err; tail -c 600 gpurun_out/bench2_$w.json; tail -3 gpurun_out/bench2_$w.err
  done
fi
