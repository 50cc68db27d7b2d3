#!/bin/bash
# Round-2 batch 27 (final code): ncu launch list of one C2 evaluation; ncu --set full of the tcgen05 update (16 launches), the leaf
# and both panel kernels; CSV exports on the box (the .ncu-rep files stay in /tmp: gpurun_out is limited to 64 MiB).
mkdir -p gpurun_out
echo "== ncu launch list (C2, second evaluation)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 170 --csv --log-file gpurun_out/b27_launches_c2.csv python scripts/time_lml.py 8192 1 ncu > gpurun_out/b27_ncu0.log 2>&1; echo "rc=$?"
echo "== ncu --set full: syrk_i8 launches 32..47"
timeout 900 ncu --set full --clock-control none -k regex:syrk_i8 -s 31 -c 16 -o /tmp/b27_syrk python scripts/time_lml.py 8192 1 ncu > gpurun_out/b27_ncu1.log 2>&1; echo "rc=$?"
ncu -i /tmp/b27_syrk.ncu-rep --page raw --csv > gpurun_out/b27_syrk_raw.csv 2>/dev/null
echo "== ncu --set full: leaf + panels"
timeout 900 ncu --set full --clock-control none -k regex:"potrf_panel|potrf_leaf" -s 130 -c 4 -o /tmp/b27_misc python scripts/time_lml.py 8192 1 ncu > gpurun_out/b27_ncu2.log 2>&1; echo "rc=$?"
ncu -i /tmp/b27_misc.ncu-rep --page raw --csv > gpurun_out/b27_misc_raw.csv 2>/dev/null
ls -la gpurun_out/b27*; du -sh gpurun_out
