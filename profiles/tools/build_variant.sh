#!/bin/sh
# build_variant.sh NAME [-DFLAG ...] -- builds an experimental libigate_dsp.so (profiles/tools/variants/lib_NAME.so,
# git-ignored, travels with gpurun) with extra macro definitions; bench it with
#   IGD_LIB_PATH=$PWD/profiles/tools/variants/lib_NAME.so python bench.py --steps 20 --no-cpu --no-e2e
set -e
HERE=$(cd "$(dirname "$0")" && pwd); ROOT=$(cd "$HERE/../.." && pwd)
NAME=$1; shift
OUT=$HERE/variants; mkdir -p $OUT/obj_$NAME
SRC=$ROOT/igate4xsoftphonedsp_b200/csrc
for u in igd_fused igd_codec igd_packet igd_walks igd_capi; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -cudart static "$@" -c -o $OUT/obj_$NAME/$u.o $SRC/$u.cu &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -cudart static -shared -o $OUT/lib_$NAME.so $OUT/obj_$NAME/*.o
rm -rf $OUT/obj_$NAME
echo built $OUT/lib_$NAME.so
