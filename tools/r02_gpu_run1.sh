#!/bin/bash
# Round 2, GPU call 1: parity (all GPU tests, kernels 6/7 un-gated), then the single-GPU
# diagnosis matrix for the headline workload: stage kernel x storage order x group schedule
# x L2 hints, DRAM bytes per launch from ncu, one full capture of the best candidate.
set -u
out=gpurun_out/r02a
mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$out/smi.txt"
echo "== tests"
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -40 | tee "$out/tests.log"
echo "== bench matrix"
b() { label=$1; shift; timeout 300 python bench.py --no-cpu --steps 10 --warmup 3 "$@" > "$out/bench_$label.json" 2> "$out/bench_$label.err"; python - "$out/bench_$label.json" "$label" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
    print(sys.argv[2], r['kernel'], 'order', d['config']['storage_order'], '%.4g ADO-steps/s' % d['value'], '%.3f ms/step' % d['ms_per_step'], 'frac %.3f' % (r['frac'] or 0), 'e2e %.4g' % d['e2e']['value'], 'rho00', d['check'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
b k3_o0 --kernel 3 --order 0
b k6_o0_static --kernel 6 --order 0 --dynsched 0
b k6_o0_dyn --kernel 6 --order 0
b k6_o2_static --kernel 6 --order 2 --dynsched 0
b k6_o2_dyn --kernel 6 --order 2
b k6_o1_dyn --kernel 6 --order 1
b k7_o0_static --kernel 7 --order 0 --dynsched 0
b k7_o0_dyn --kernel 7 --order 0
b k7_o2_static --kernel 7 --order 2 --dynsched 0
b k7_o2_dyn --kernel 7 --order 2
b k7_o1_dyn --kernel 7 --order 1
b k7_o2_dyn_pref --kernel 7 --order 2 --prefetch 1
export PYQED_HEOM_LIB=$PWD/pyqed_b200/lib/libpyqed_heom_h1.so
b k7_o2_dyn_h1 --kernel 7 --order 2
b k6_o2_dyn_h1 --kernel 6 --order 2
export PYQED_HEOM_LIB=$PWD/pyqed_b200/lib/libpyqed_heom_h2.so
b k7_o2_dyn_h2 --kernel 7 --order 2
b k6_o2_dyn_h2 --kernel 6 --order 2
unset PYQED_HEOM_LIB
echo "== DRAM bytes per launch (ncu, 4 launches of one step each)"
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum
n() { label=$1; shift; timeout 400 ncu --metrics $M --clock-control none -k regex:stage_rows -s 16 -c 4 --csv --log-file "$out/ncu_$label.csv" python bench.py --no-cpu --steps 2 --warmup 3 "$@" > "$out/ncu_$label.log" 2>&1; python - "$out/ncu_$label.csv" "$label" <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10 and r[0].isdigit()]
acc={}
for r in rows:
    acc.setdefault(r[0],{})[r[-3]]=r[-1]
for k,v in acc.items():
    print(sys.argv[2], k, {m.split('__')[-1][:22]:x for m,x in v.items()})
PY
}
n k3_o0 --kernel 3 --order 0
n k6_o0_dyn --kernel 6 --order 0
n k6_o2_static --kernel 6 --order 2 --dynsched 0
n k6_o2_dyn --kernel 6 --order 2
n k7_o0_dyn --kernel 7 --order 0
n k7_o2_static --kernel 7 --order 2 --dynsched 0
n k7_o2_dyn --kernel 7 --order 2
export PYQED_HEOM_LIB=$PWD/pyqed_b200/lib/libpyqed_heom_h1.so
n k7_o2_dyn_h1 --kernel 7 --order 2
unset PYQED_HEOM_LIB
echo "== full capture: kernel 7, order 2, dynamic schedule (middle, middle, last stage)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage_rows_sym_kernel -s 13 -c 3 \
    -o "$out/k7_o2_dyn_full" python bench.py --no-cpu --kernel 7 --order 2 --steps 2 --warmup 3 > "$out/ncu_full.log" 2>&1
ls -la "$out" | tail -50
