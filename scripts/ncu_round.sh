#!/bin/bash
# ncu --set full captures of the four hot kernels (one launch each) on one GPR LML evaluation
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:kbuild_fast_kernel -c 1 -o gpurun_out/ncu_kbuild python scripts/one_lml.py 8192 1 > gpurun_out/ncu1.log 2>&1
$NCU -k regex:syrk_i8_kernel -s 7 -c 1 -o gpurun_out/ncu_syrk python scripts/one_lml.py 8192 1 > gpurun_out/ncu2.log 2>&1
$NCU -k regex:potrf_leaf_kernel -s 3 -c 1 -o gpurun_out/ncu_leaf python scripts/one_lml.py 8192 1 > gpurun_out/ncu3.log 2>&1
$NCU -k regex:gemm_dmma_kernel -s 20 -c 1 -o gpurun_out/ncu_dmma python scripts/one_lml.py 8192 1 > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep
