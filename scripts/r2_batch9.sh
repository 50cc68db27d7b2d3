#!/bin/bash
# Round-2 batch 9: critical-row CTA mapping of the fused panel: tests + timing.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/b9_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/b9_pytest.log | cut -c1-200
run() { env "$@" timeout 300 python scripts/time_lml.py 8192 10 "$*" 2>&1 | tail -1 | tee -a gpurun_out/b9_ab.txt; }
run X=default
run GPK_PANEL_FUSE=0
run X=default2
for n in 1000 2048 4096 5000; do timeout 300 python scripts/time_lml.py $n 10 "N=$n" 2>&1 | tail -1 | tee -a gpurun_out/b9_ab.txt; done
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 210 --csv --log-file gpurun_out/b9_launches_c2.csv python scripts/time_lml.py 8192 1 ncu > gpurun_out/b9_ncu0.log 2>&1; echo "rc=$?"
