#!/bin/bash
# tools/build_variant.sh NAME "-DFOO=1 ..." : an A/B build of the library with extra defines ->
# osep_path_planner_b200/libosep_skel_NAME.so (select it with OSEP_SKEL_LIB=...)
set -e
cd "$(dirname "$0")/../osep_path_planner_b200/csrc"
NAME=$1; DEFS=$2
OUT=/tmp/osep_variant_$NAME; mkdir -p $OUT
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC,-ffp-contract=off,-O3 -Xptxas -v"
for f in nscale mscale drosa graph api map; do
  nvcc $FLAGS $DEFS -c $f.cu -o $OUT/$f.o 2> $OUT/$f.ptxas.log &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libosep_skel_$NAME.so $OUT/*.o -lcudart
echo built ../libosep_skel_$NAME.so
