#!/bin/bash
# First GPU round: smoke, parity tests, bench, ncu launch list.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" > gpurun_out/host.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== sanitizer (smoke)"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py --smoke > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -4 gpurun_out/sanitizer.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench c1"; timeout 300 python bench.py --workload gpr_c1 --steps 20 > gpurun_out/bench_c1.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_c1.json
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu rc=$?"
