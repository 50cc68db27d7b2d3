#!/bin/bash
# Round-2 batch 5: batched epilogue of syrk_i8; gradient / edge tests.
mkdir -p gpurun_out
run() { env "$@" timeout 300 python scripts/time_lml.py 8192 8 "$*" 2>&1 | tail -1 | tee -a gpurun_out/b5_ab.txt; }
run X=default
run GPK_TC_STATIC=0
run GPK_PANEL_FUSE=0
run GPK_TC_MIN_K=256
run GPK_TC_MIN_K=128
run GPK_TC_CAT=0
run GPK_TC_CAT=0 GPK_TC_A_TMEM=1
run GPK_TC_SLICES=8
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/b5_pytest.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/b5_pytest.log
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 190 --csv --log-file gpurun_out/b5_launches.csv python scripts/time_lml.py 8192 1 ncu > gpurun_out/b5_ncu.log 2>&1; echo "ncu rc=$?"
