#!/bin/bash
# --set full captures of kernel 9 on config 4 (every launch of a short bench run; the long ones are the timed region)
set -u
out=gpurun_out/${1:-k9ncu}
mkdir -p "$out"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:stage_dataflow_tma -c 5 -f -o "$out/k9_full" python bench.py --no-cpu --workload polariton32_K4_L6 --steps 300 --warmup 1 > "$out/ncu_k9.log" 2>&1
tail -3 "$out/ncu_k9.log"
ncu -i "$out/k9_full.ncu-rep" --page raw --csv --metrics gpu__time_duration.sum,smsp__inst_executed.sum 2>/dev/null | tail -6
ls -la "$out"
