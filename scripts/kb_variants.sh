#!/bin/bash
# Builds libgpk variants with parts of the fast K-build kernel disabled (KF_VARIANT bit mask) to see where its time
# goes.  Run here (cross-compile), then `scripts/kb_variants.sh run` on the GPU box.
cd "$(dirname "$0")/.."
VDIR=gpflow_b200/build/variants
if [ "$1" != "run" ]; then
  mkdir -p $VDIR
  python gpflow_b200/build.py
  for v in 0 1 4 7; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
      -DGPK_BUILD -DKF_VARIANT=$v -cudart static -c gpflow_b200/csrc/kbuild.cu -o $VDIR/kbuild_v$v.o &
  done
  wait
  for v in 0 1 4 7; do
    objs=$(ls gpflow_b200/build/*.o | grep -v kbuild.o)
    nvcc -shared -gencode arch=compute_100a,code=sm_100a -cudart static -o $VDIR/libgpk_v$v.so $objs $VDIR/kbuild_v$v.o
  done
  ls -la $VDIR/*.so
else
  for v in 0 1 4 7; do
    GPFLOW_B200_LIB=$PWD/$VDIR/libgpk_v$v.so python scripts/kb_time.py $v
  done
fi
