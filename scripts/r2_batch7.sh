#!/bin/bash
# Round-2 batch 7 (2 GPUs): NCCL parity test of the sharded SVGP evaluations, the fixed tests, 2-GPU bench lines.
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest (multirank + fixed + c3 predict)"
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_grad.py tests/test_gpu_widen.py tests/test_gpu_widen2.py "tests/test_gpu_models.py::test_c3_full_size_sgpr_predict_golden" "tests/test_gpu_models.py::test_c4_full_size_svgp_elbo_golden" -m gpu -q --timeout 600 > gpurun_out/b7_pytest.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/b7_pytest.log | cut -c1-200
echo "== svgp batched A/B (1 GPU)"
for b in 1 0; do GPK_SVGP_BATCHED=$b timeout 600 python bench.py --workload svgp_c4 --steps 20 --no-svgp 2> gpurun_out/b7_svgp_$b.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('batched=$b', d['ms_per_step'], d['value'], d['objective_vs_cpu_rel_err'], d['roofline']['frac'], d['kernel_classes'])"; done
echo "== bench --gpus 2 (C2 + SVGP modes)"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/b7_bench2.json 2> gpurun_out/b7_bench2.err; echo "rc=$?"; tail -3 gpurun_out/b7_bench2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/b7_bench2.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], d['step_stats'])
for k,v in d['svgp_c4'].items():
    if isinstance(v,dict): print(k, v['evals_per_s'], v['ms_per_step'], v.get('sum_of_shares_vs_full_rel_err'), v['step_stats']['per_rank_median_ms'])
PY
