#!/bin/bash
# build (cross-compiles without a GPU) and, with a GPU present, run the tcgen05.mma rate microbenchmark
cd "$(dirname "$0")" && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o mma_rate.bin mma_rate.cu && \
  (nvidia-smi -L > /dev/null 2>&1 && ./mma_rate.bin || echo "built tools/mma_rate.bin (no GPU here)")
