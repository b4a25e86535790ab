#!/bin/bash
# Build an A/B variant of the library: tools/build_variant.sh NAME -DMACRO[=V] ... -- file.cu ...
# -> adrt_b200/csrc/build_ab/libadrt_b200_NAME.so (the named sources recompiled with the macros, the other
# objects taken from the default build); run with ADRT_B200_LIB=$PWD/adrt_b200/csrc/build_ab/libadrt_b200_NAME.so
set -e
cd "$(dirname "$0")/../adrt_b200/csrc"
name=$1; shift
defs=()
while [ "$1" != "--" ]; do defs+=("$1"); shift; done
shift
mkdir -p build_ab
objs=()
for o in api host_api step_kernels fused_adrt stream_adrt stage_adrt iadrt_fused; do
  if [[ " $* " == *" $o.cu "* ]]; then
    /usr/local/cuda/bin/nvcc -ccbin /usr/bin/g++ -O3 -std=c++17 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a \
      -Xcompiler -fPIC,-ffp-contract=off "${defs[@]}" -c $o.cu -o build_ab/${o}_$name.o
    objs+=(build_ab/${o}_$name.o)
  else
    objs+=(build/$o.o)
  fi
done
/usr/local/cuda/bin/nvcc -ccbin /usr/bin/g++ -shared -gencode arch=compute_100a,code=sm_100a -o build_ab/libadrt_b200_$name.so "${objs[@]}" -lpthread
echo build_ab/libadrt_b200_$name.so
