#!/bin/bash
# Round-2 batch 3: why is syrk_i8 slow with the global plane store?  A/B of layout / MMA form + ncu of two launches.
mkdir -p gpurun_out
run() { env "$@" timeout 300 python scripts/time_lml.py 8192 6 "$*" 2>&1 | tail -1 | tee -a gpurun_out/b3_ab.txt; }
run X=default
run GPK_TC_RECT=1
run GPK_TC_CAT=0
run GPK_TC_CAT=0 GPK_TC_A_TMEM=1
run GPK_TC_CLUSTER=1
run GPK_LOOKAHEAD=0
echo "== ncu syrk (launch 1 = K512, launch 8 = K4096 of the first captured evaluation)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:syrk_i8 -s 15 -c 8 -o gpurun_out/b3_syrk python scripts/time_lml.py 8192 1 ncu > gpurun_out/b3_ncu.log 2>&1; echo "ncu rc=$?"
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/b3_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/b3_pytest.log
echo "== bench (no svgp)"; timeout 600 python bench.py --steps 5 --warmup 3 --no-svgp > gpurun_out/b3_bench.json 2> gpurun_out/b3_bench.err; echo "rc=$?"; tail -3 gpurun_out/b3_bench.err; head -c 1500 gpurun_out/b3_bench.json
