#!/bin/bash
# Builds libgpk variants with the three 32x32 diagonal-Cholesky formulations (GPK_CHOL32_VARIANT 0/1/2) and, with
# `run`, times the leaf phases and the C2 step for each on the GPU box.
cd "$(dirname "$0")/.."
VDIR=gpflow_b200/build/variants
if [ "$1" != "run" ]; then
  mkdir -p $VDIR
  python gpflow_b200/build.py
  for v in 0 2; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
      -DGPK_BUILD -DGPK_CHOL32_VARIANT=$v -cudart static -c gpflow_b200/csrc/potrf.cu -o $VDIR/potrf_v$v.o &
  done
  wait
  for v in 0 2; do
    objs=$(ls gpflow_b200/build/*.o | grep -v "/potrf.o")
    nvcc -shared -gencode arch=compute_100a,code=sm_100a -cudart static -o $VDIR/libgpk_c$v.so $objs $VDIR/potrf_v$v.o
  done
  ls -la $VDIR/*.so
else
  for v in 1 0 2; do
    L=$PWD/$VDIR/libgpk_c$v.so; [ $v = 1 ] && L=$PWD/gpflow_b200/libgpk.so
    echo "== chol32 variant $v"
    GPFLOW_B200_LIB=$L timeout 120 python scripts/leaf_timing.py 2>&1 | tail -2
    GPFLOW_B200_LIB=$L timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('step ms', round(d['ms_per_step'],3))"
  done
fi
