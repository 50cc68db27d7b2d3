#!/bin/bash
# Round-2 batch 15: chain shortening (no counter memsets, extra rows sliced by the panel kernel, stream assignment): GPU suite + timings.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/b15_pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/b15_pytest.log
run() { env "$@" timeout 300 python scripts/time_lml.py ${N:-8192} 10 "$*" 2>&1 | tail -1 | tee -a gpurun_out/b15_ab.txt; }
run X=default
run X=default2
run GPK_TC_SLICES=7
N=4096 run X=default
N=2048 run X=default
N=16384 run X=default
