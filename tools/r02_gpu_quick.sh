#!/bin/bash
# quick single-GPU check of a stage-kernel change: kernel 6/7 parity tests, bench, per-launch DRAM / shared-memory counters
set -u
out=gpurun_out/${1:-quick}
mkdir -p "$out"
timeout 900 python -m pytest tests/test_gpu_kernel6.py -m gpu -q -x 2>&1 | tail -5 | tee "$out/tests.log"
b() { label=$1; shift; timeout 300 python bench.py --no-cpu --steps 10 --warmup 3 "$@" > "$out/bench_$label.json" 2> "$out/bench_$label.err"; python - "$out/bench_$label.json" "$label" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
    print(sys.argv[2], r['kernel'], 'order', d['config']['storage_order'], '%.4g ADO-steps/s' % d['value'], '%.3f ms/step' % d['ms_per_step'], 'frac %.3f' % (r['frac'] or 0), 'e2e %.4g' % d['e2e']['value'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e, open(sys.argv[1].replace('.json','.err')).read()[-600:])
PY
}
b k7_o2 --kernel 7 --order 2
b k6_o2 --kernel 6 --order 2
b auto_o2 --order 2
b auto_o0
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active
n() { label=$1; shift; timeout 400 ncu --metrics $M --clock-control none -k regex:stage_rows -s 16 -c 4 --csv --log-file "$out/ncu_$label.csv" python bench.py --no-cpu --steps 2 --warmup 3 "$@" > "$out/ncu_$label.log" 2>&1; python - "$out/ncu_$label.csv" "$label" <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10 and r[0].isdigit()]
acc={}
for r in rows:
    acc.setdefault(r[0],{})[r[-3]]=r[-1]
for k,v in acc.items():
    print(sys.argv[2], k, {m.split('__')[-1][:26]:x for m,x in v.items()})
PY
}
n k7_o2 --kernel 7 --order 2
n k6_o2 --kernel 6 --order 2
