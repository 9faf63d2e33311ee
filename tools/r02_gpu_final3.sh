#!/bin/bash
# Round 2, last check of the final build on one GPU: GPU tests, smoke, the default bench as the driver runs it,
# config 4, ncu capture of kernel 9
set -u
out=gpurun_out/${1:-r02final3}
mkdir -p "$out"
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -rfs 2>&1 | tail -8 | tee "$out/tests.log"
echo "== smoke"
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -9 | tee "$out/smoke.log"
echo "== config 4"
timeout 200 python bench.py --no-cpu --workload polariton32_K4_L6 --warmup 3 --steps 3000 > "$out/bench_cfg4_k9.json" 2> "$out/bench_cfg4_k9.err"
python - "$out/bench_cfg4_k9.json" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1]); r=d['roofline']
print(r['kernel'], '%.4g ADO-steps/s' % d['value'], '%.2f us/step' % (1e3*d['ms_per_step']), 'e2e %.4g' % d['e2e']['value'])
PY
echo "== bench (as the driver runs it)"
s=$(date +%s); timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > "$out/bench_default.json" 2> "$out/bench_default.err"; e=$(date +%s); echo "wall $((e-s)) s"
python - "$out/bench_default.json" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
r=d['roofline']; c=d['check']
print('value %.4g ms/step %.3f frac %.3f traffic %s e2e %.4g launches %d clocks %s' % (d['value'], d['ms_per_step'], r['frac'], r['traffic'], d['e2e']['value'], d['gpu_launches'], d['clocks']))
print('check', {k:v for k,v in c.items() if k!='rho_sys_final'})
print('other', d.get('other_workloads'))
PY
echo "== ncu kernel 9"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:stage_dataflow_tma -s 1 -c 1 -f -o "$out/k9_full" python bench.py --no-cpu --workload polariton32_K4_L6 --steps 300 --warmup 1 > "$out/ncu_k9.log" 2>&1
tail -2 "$out/ncu_k9.log"
