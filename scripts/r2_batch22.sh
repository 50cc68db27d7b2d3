#!/bin/bash
# Round-2 batch 22: flag hops with the grid-size guard: large-N runs, the full GPU suite, timings.
mkdir -p gpurun_out
for n in 16384 12288 9856; do timeout 300 python scripts/time_lml.py $n 5 n$n 2>&1 | tail -1; done
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/b22_pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/b22_pytest.log
run() { env "$@" timeout 300 python scripts/time_lml.py ${N:-8192} 10 "$*" 2>&1 | tail -1 | tee -a gpurun_out/b22_ab.txt; }
run X=default
run GPK_FLAG_HOPS=0
