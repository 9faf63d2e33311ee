#!/bin/bash
# Round 2, GPU call: full GPU test suite, FP64 microbenchmark, configs 4 and 5, headline
set -u
out=gpurun_out/${1:-r02g}
mkdir -p "$out"
echo "== fp64 peaks"
bash tools/fp64_peaks.sh | tee "$out/fp64_peaks.json"
echo "== tests"
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -25 | tee "$out/tests.log"
b() { label=$1; shift; timeout 300 python bench.py --no-cpu --warmup 3 "$@" > "$out/bench_$label.json" 2> "$out/bench_$label.err"; python - "$out/bench_$label.json" "$label" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1]); r=d['roofline']
    print(sys.argv[2], r['kernel'], d['config']['workload'], '%.4g ADO-steps/s' % d['value'], '%.4f ms/step' % d['ms_per_step'], 'frac %.3f' % (r['frac'] or 0), 'e2e %.4g' % d['e2e']['value'], 'launches', d['gpu_launches'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e, open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
}
b cfg4_k8 --workload polariton32_K4_L6 --steps 500
b cfg4_k2 --workload polariton32_K4_L6 --steps 500 --kernel 2
b cfg5_b64 --workload aggregate7_K6_L6 --batch 64 --steps 700
b cfg5_b1 --workload aggregate7_K6_L6 --steps 700
b cfg2 --workload fmo7_K7_L4 --steps 1000
b cfg1 --workload spin_boson_K2_L10 --steps 1000
b headline --steps 20
echo "== ncu: kernel 8 on config 4"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stage_dataflow -c 1 -o "$out/k8_cfg4" python bench.py --no-cpu --workload polariton32_K4_L6 --steps 50 --warmup 3 > "$out/ncu_k8.log" 2>&1
ncu -i "$out/k8_cfg4.ncu-rep" --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
for d in rows[2:3]:
    for w in ['gpu__time_duration.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__grid_size','launch__registers_per_thread']:
        if w in hdr: print(w, d[hdr.index(w)])
" | tee "$out/k8_cfg4_summary.txt"
ls "$out"
