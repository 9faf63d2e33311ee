#!/bin/bash
# kernel 9 stage breakdown (clock64 marks, -DDF9_PROFILE build) + parity + bench
set -u
out=gpurun_out/${1:-k9p}
mkdir -p "$out"
# the profiling build (clock64 marks per stage phase); built here if it did not travel with the snapshot
[ -f pyqed_b200/lib/libpyqed_heom_prof.so ] || python -c "
import sys; sys.path.insert(0, '.')
from pyqed_b200 import build
build.build_extension(defines=['DF9_PROFILE'], out='$PWD/pyqed_b200/lib/libpyqed_heom_prof.so')"
PYQED_HEOM_LIB=$PWD/pyqed_b200/lib/libpyqed_heom_prof.so timeout 200 python bench.py --no-cpu --workload polariton32_K4_L6 --warmup 1 --steps 2000 2>&1 | grep "^df9" | sort | uniq -c | tee "$out/prof.txt"
bash tools/r02_gpu_k9.sh "${1:-k9p}"
