#!/bin/bash
# one --set full capture of kernel 9 on config 4 (source counters included)
set -u
out=gpurun_out/${1:-k9ncu}
mkdir -p "$out"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:stage_dataflow_tma -c 1 -f -o "$out/k9_full" python bench.py --no-cpu --workload polariton32_K4_L6 --steps 300 --warmup 1 > "$out/ncu_k9.log" 2>&1
tail -3 "$out/ncu_k9.log"
ls -la "$out"
