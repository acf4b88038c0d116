#!/bin/bash
# build_variants/<name>.so = the library with the fused-solve translation units recompiled with extra flags
# usage: scripts/build_variant.sh NAME "-DTODE_FUSED_MINB=5 ..."
set -e
NAME=$1; FLAGS=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/torchode_b200/csrc
OUT=$ROOT/build_variants/$NAME
mkdir -p $OUT
CXX=$(test -x /usr/bin/g++ && echo /usr/bin/g++ || echo g++)
NVF="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -ccbin $CXX -Xcompiler -fPIC -Xcompiler -O2"
pids=()
for f in ${VARIANT_TUS:-fused_f64f64 fused_f32f32}; do
  nvcc $NVF $FLAGS -Xptxas -v -c $SRC/$f.cu -o $OUT/$f.o 2> $OUT/$f.ptxas.log &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
OBJS=""
for o in $SRC/build/*.o; do
  b=$(basename $o)
  if [ -f $OUT/$b ]; then OBJS="$OBJS $OUT/$b"; else OBJS="$OBJS $o"; fi
done
nvcc -shared -o $ROOT/build_variants/$NAME.so $OBJS -cudart static 2>/dev/null
echo "built build_variants/$NAME.so"
