#!/bin/bash
# A/B of two builds of kernel 9 on config 4 (device-timed, 3 runs each) + the dataflow parity tests on build B
set -u
out=gpurun_out/${1:-k9ab}; mkdir -p "$out"
B=${2:-libpyqed_heom_b.so}
PYQED_HEOM_LIB=$PWD/pyqed_b200/lib/$B timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dataflow" 2>&1 | tail -3 | tee "$out/tests_b.log"
for i in 1 2 3; do
  for lib in libpyqed_heom.so $B; do
    PYQED_HEOM_LIB=$PWD/pyqed_b200/lib/$lib timeout 200 python bench.py --no-cpu --workload polariton32_K4_L6 --warmup 3 --steps 3000 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('$lib', d['roofline']['kernel'], '%.2f us/step' % (1e3*d['ms_per_step']), '%.4g' % d['value'])" | tee -a "$out/ab.txt"
  done
done
