#!/bin/bash
# Round-2 batch 2: static digit planes + fused panel/update: parity tests, A/B timings, launch list.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/b2_pytest.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/b2_pytest.log
echo "== A/B"
run() { env "$@" timeout 300 python scripts/time_lml.py 8192 10 "$*" 2>&1 | tail -1 | tee -a gpurun_out/b2_ab.txt; }
run X=default
run GPK_PANEL_FUSE=0
run GPK_TC_STATIC=0
run GPK_TC_STATIC=0 GPK_PANEL_FUSE=0
run GPK_TC_MIN_K=256
run GPK_TC_MIN_K=128
run GPK_TC_MIN_K=256 GPK_PANEL_FUSE=0
run GPK_TC_SLICES=8
run GPK_FP64_ENGINE=dmma
echo "== other sizes (default)"
for n in 1000 2048 4096 5000; do timeout 300 python scripts/time_lml.py $n 10 "N=$n" 2>&1 | tail -1 | tee -a gpurun_out/b2_ab.txt; done
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 200 --csv --log-file gpurun_out/b2_launches.csv python scripts/time_lml.py 8192 1 ncu > gpurun_out/b2_ncu.log 2>&1; echo "ncu rc=$?"
