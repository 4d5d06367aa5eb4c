#!/bin/bash
# Builds experimental variants of libpfcu.so with extra -D flags: tools/variants.sh name "-DFOO=1" ...
# A name starting with all_ applies the flags to every source file, otherwise only to the raster kernels.
# Output: pathfinder-cpp_b200/lib/libpfcu_<name>.so (git-ignored; travels to the GPU box).
set -e
cd "$(dirname "$0")/../pathfinder-cpp_b200"
make -s
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  if [[ $name == all_* ]]; then
    $NV $flags -fmad=false -c csrc/pfcu_geom.cu -o build/pfcu_geom_$name.o &
    $NV $flags -c csrc/pfcu_tiles.cu -o build/pfcu_tiles_$name.o &
    $NV $flags -c csrc/pfcu_api.cu -o build/pfcu_api_$name.o &
    $NV $flags -c csrc/pfcu_raster.cu -o build/pfcu_raster_$name.o &
    wait
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o lib/libpfcu_$name.so build/pfcu_geom_$name.o build/pfcu_stroke.o build/pfcu_tiles_$name.o build/pfcu_api_$name.o build/pfcu_raster_$name.o
  else
    $NV $flags -c csrc/pfcu_raster.cu -o build/pfcu_raster_$name.o
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o lib/libpfcu_$name.so build/pfcu_geom.o build/pfcu_stroke.o build/pfcu_tiles.o build/pfcu_api.o build/pfcu_raster_$name.o
  fi
done
