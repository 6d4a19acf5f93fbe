#!/bin/bash
# usage: scratch/build_variant.sh NAME FILE.cu "-DFOO=1 ..."   -> scratch/variants/libast_NAME.so (other objects from csrc/build)
set -e
cd "$(dirname "$0")/../artist_style_transfer_b200/csrc"
make -s
NAME=$1; FILE=$2; DEFS=$3
mkdir -p build_var ../../scratch/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $DEFS -c $FILE -o build_var/${NAME}_${FILE%.cu}.o
OBJS=$(ls build/*.o | grep -v "build/${FILE%.cu}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../scratch/variants/libast_${NAME}.so $OBJS build_var/${NAME}_${FILE%.cu}.o -Xlinker --version-script=exports.map
echo built scratch/variants/libast_${NAME}.so
