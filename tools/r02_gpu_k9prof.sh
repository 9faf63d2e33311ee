#!/bin/bash
# kernel 9 stage breakdown (clock64 marks, -DDF9_PROFILE build) + parity + bench
set -u
out=gpurun_out/${1:-k9p}
mkdir -p "$out"
PYQED_HEOM_LIB=$PWD/pyqed_b200/lib/libpyqed_heom_prof.so timeout 200 python bench.py --no-cpu --workload polariton32_K4_L6 --warmup 1 --steps 2000 2>&1 | grep "^df9" | sort | uniq -c | tee "$out/prof.txt"
bash tools/r02_gpu_k9.sh "${1:-k9p}"
