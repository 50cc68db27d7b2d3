#!/bin/bash
# Round-2 batch 1: tcgen05 MMA issue-rate probe, leaf phase timing, A/B of the concatenated-N MMA variants on C2.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/b1_smi.txt 2>&1
echo "== mb_mma"; timeout 120 ./scripts/mb_mma.bin 2>&1 | tee gpurun_out/b1_mb_mma.txt
echo "== leaf timing"; timeout 120 python scripts/leaf_timing.py 2>&1 | tee gpurun_out/b1_leaf.txt
echo "== A/B"
for cat in 0 1; do for ts in 1 0; do
  GPK_TC_CAT=$cat GPK_TC_A_TMEM=$ts timeout 300 python scripts/time_lml.py 8192 10 "cat=$cat ts=$ts" 2>&1 | tail -1 | tee -a gpurun_out/b1_ab.txt
done; done
GPK_TC_CAT=1 GPK_TC_A_TMEM=0 GPK_TC_CLUSTER=1 timeout 300 python scripts/time_lml.py 8192 10 "cat=1 ts=0 cl=1" 2>&1 | tail -1 | tee -a gpurun_out/b1_ab.txt
GPK_TC_CAT=1 GPK_TC_A_TMEM=1 GPK_TC_MIN_K=256 timeout 300 python scripts/time_lml.py 8192 10 "cat=1 ts=1 mink=256" 2>&1 | tail -1 | tee -a gpurun_out/b1_ab.txt
echo "== tc tests (cat=1 ts=1 / ts=0)"
GPK_TC_CAT=1 GPK_TC_A_TMEM=1 timeout 600 python -m pytest tests/test_gpu_tc.py -q -x 2>&1 | tail -3
GPK_TC_CAT=1 GPK_TC_A_TMEM=0 timeout 600 python -m pytest tests/test_gpu_tc.py -q -x 2>&1 | tail -3
