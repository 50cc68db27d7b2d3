#!/bin/bash
# Round-2 batch 10: 5-stage pipeline; ncu --set full captures exported to CSV on the box (the .ncu-rep files stay there).
mkdir -p gpurun_out
echo "== tests (tc + kernels + models subset)"; timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_kernels.py tests/test_gpu_edge.py -m gpu -q --timeout 600 2>&1 | tail -3
run() { env "$@" timeout 300 python scripts/time_lml.py 8192 10 "$*" 2>&1 | tail -1 | tee -a gpurun_out/b10_ab.txt; }
run X=default
run X=default2
echo "== ncu --set full: syrk_i8 launches 32..47 of the run (second evaluation: K = 256,512,256,1024,...,4096)"
timeout 900 ncu --set full --clock-control none -k regex:syrk_i8 -s 31 -c 16 -o /tmp/b10_syrk python scripts/time_lml.py 8192 1 ncu > gpurun_out/b10_ncu1.log 2>&1; echo "rc=$?"
ncu -i /tmp/b10_syrk.ncu-rep --page raw --csv > gpurun_out/b10_syrk_raw.csv 2>/dev/null
echo "== ncu --set full: kbuild_fast, panel (fused + plain), leaf"
timeout 900 ncu --set full --clock-control none -k regex:"kbuild_fast|potrf_panel|potrf_leaf" -s 130 -c 5 -o /tmp/b10_misc python scripts/time_lml.py 8192 1 ncu > gpurun_out/b10_ncu2.log 2>&1; echo "rc=$?"
ncu -i /tmp/b10_misc.ncu-rep --page raw --csv > gpurun_out/b10_misc_raw.csv 2>/dev/null
echo "== ncu --set full: batched tf32 (SVGP)"
timeout 900 ncu --set full --clock-control none -k regex:gemm_tf32 -s 30 -c 2 -o /tmp/b10_tf32 python bench.py --workload svgp_c4 --steps 3 --no-svgp > gpurun_out/b10_ncu3.log 2>&1; echo "rc=$?"
ncu -i /tmp/b10_tf32.ncu-rep --page raw --csv > gpurun_out/b10_tf32_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -12; du -sh gpurun_out
