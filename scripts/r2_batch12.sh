#!/bin/bash
# Round-2 batch 12: where does a K = 256 syrk_i8 launch spend its time?  Source-level sampling of one launch.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:syrk_i8 -s 31 -c 1 -o /tmp/b12_k256 python scripts/time_lml.py 8192 1 ncu > gpurun_out/b12_ncu.log 2>&1; echo "rc=$?"
ncu -i /tmp/b12_k256.ncu-rep --page source --csv > gpurun_out/b12_k256_source.csv 2>/dev/null
ls -la gpurun_out/b12_k256_source.csv
