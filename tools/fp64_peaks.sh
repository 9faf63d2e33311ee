#!/bin/bash
# builds (here or on the GPU box) and runs the FP64 microbenchmark; prints one JSON line
set -e
cd "$(dirname "$0")"
[ -x ./fp64_peaks ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
./fp64_peaks
