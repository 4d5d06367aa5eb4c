#!/bin/bash
# Builds experimental variants of libpfcu.so with extra -D flags for the raster kernels: tools/variants.sh name "-DFOO=1" ...
# Output: pathfinder-cpp_b200/lib/libpfcu_<name>.so (git-ignored; travels to the GPU box).
set -e
cd "$(dirname "$0")/../pathfinder-cpp_b200"
make -s
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $flags -c csrc/pfcu_raster.cu -o build/pfcu_raster_$name.o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o lib/libpfcu_$name.so build/pfcu_geom.o build/pfcu_tiles.o build/pfcu_api.o build/pfcu_raster_$name.o
done
