#!/bin/bash
# ncu --set full captures (one launch each): tf32 tcgen05 GEMM inside an SVGP evaluation, Cholesky panel kernel and slim leaf
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:gemm_tf32_kernel -s 10 -c 1 -o gpurun_out/ncu_tf32 python scripts/one_eval.py svgp_c4 1 > gpurun_out/ncu5.log 2>&1
$NCU -k regex:potrf_panel_kernel -s 3 -c 1 -o gpurun_out/ncu_panel python scripts/one_lml.py 8192 1 > gpurun_out/ncu6.log 2>&1
$NCU -k regex:potrf_leaf_kernel -s 3 -c 1 -o gpurun_out/ncu_leaf python scripts/one_lml.py 8192 1 > gpurun_out/ncu7.log 2>&1
ls -la gpurun_out/*.ncu-rep
