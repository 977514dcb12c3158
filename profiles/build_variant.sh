#!/bin/bash
# Builds a variant of the library with extra -D flags for a kernel experiment:
#   profiles/build_variant.sh NAME "-DFGS_ES_XORCMP"   ->  profiles/variants/libflexgsea_b200_NAME.so
# (git-ignored, travels to the GPU box; select it with FLEXGSEA_B200_LIB=<path>)
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2
out=profiles/variants; obj=/tmp/fgs_variant_$name
mkdir -p $out $obj
for f in flexgsea-r_b200/csrc/*.cu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude $flags -c $f -o $obj/$(basename $f .cu).o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libflexgsea_b200_$name.so $obj/*.o -lcudart -ldl
echo built $out/libflexgsea_b200_$name.so
