#!/bin/bash
mkdir -p gpurun_out
for hops in 1 0; do
  echo "== N=16384 GPK_FLAG_HOPS=$hops"; GPK_FLAG_HOPS=$hops timeout 300 python scripts/time_lml.py 16384 5 hops$hops 2>&1 | tail -8
done
echo "== N=12288"; timeout 300 python scripts/time_lml.py 12288 5 n12288 2>&1 | tail -4
