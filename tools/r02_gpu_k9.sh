#!/bin/bash
# kernel 9 (heom_dataflow_tma.cuh) on one GPU: parity tests of kernels 8 / 9, config-4 bench with both, launch counters
set -u
out=gpurun_out/${1:-k9}
mkdir -p "$out"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dataflow" 2>&1 | tail -6 | tee "$out/tests.log"
b() { label=$1; shift; timeout 200 python bench.py --no-cpu --workload polariton32_K4_L6 --warmup 3 "$@" > "$out/bench_$label.json" 2> "$out/bench_$label.err"; python - "$out/bench_$label.json" "$label" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1]); r=d['roofline']
    print(sys.argv[2], r['kernel'], '%.4g ADO-steps/s' % d['value'], '%.2f us/step' % (1e3*d['ms_per_step']), 'e2e %.4g' % d['e2e']['value'], 'check', {k:v for k,v in d['check'].items() if k!='rho_sys_final'})
except Exception as e:
    print(sys.argv[2], 'FAILED', e, open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
}
b k9 --steps 2000
b k8 --steps 2000 --kernel 8
if [ "${2:-}" = "ncu" ]; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:stage_dataflow_tma -c 1 -o "$out/k9_full" python bench.py --no-cpu --workload polariton32_K4_L6 --steps 300 --warmup 1 > "$out/ncu_k9.log" 2>&1
  ncu -i "$out/k9_full.ncu-rep" --page raw --csv > "$out/k9_full_raw.csv" 2>/dev/null
  ls -la "$out"
fi
