#!/bin/bash
mkdir -p gpurun_out
for w in svgp_c4 sgpr_c3; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$w.csv python scripts/one_eval.py $w 1 > gpurun_out/l_$w.log 2>&1
tail -1 gpurun_out/l_$w.log
done
