#!/bin/bash
# Round 2 final check on one GPU: what the driver runs (GPU tests, smoke, both bench arms) + the 1-GPU rho_sys references
set -u
out=gpurun_out/${1:-r02final}
mkdir -p "$out"
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -rfs 2>&1 | tail -14 | tee "$out/tests.log"
echo "== smoke"
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -8 | tee "$out/smoke.log"
echo "== rho_sys references (1 GPU) for several step counts"
cp profiles/r02_rho_sys_n1.json "$out/rho_sys_n1.json" 2>/dev/null
for k in 5 10 15 20 25 30 40 50; do
  timeout 300 python bench.py --no-cpu --steps $k --warmup 3 --save-rho-ref "$out/rho_sys_n1.json" > /dev/null 2> "$out/rho_$k.err" || echo "rho ref $k failed"
done
cp "$out/rho_sys_n1.json" profiles/r02_rho_sys_n1.json
echo "== bench --impl reference (as the driver runs it)"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > "$out/bench_reference.json" 2> "$out/bench_reference.err"; tail -c 900 "$out/bench_reference.json"; echo
echo "== bench (as the driver runs it)"
s=$(date +%s); timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > "$out/bench_default.json" 2> "$out/bench_default.err"; e=$(date +%s); echo "wall $((e-s)) s"
python - "$out/bench_default.json" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
r=d['roofline']; c=d['check']
print('value %.4g ms/step %.3f frac %.3f traffic %s e2e %.4g launches %d clocks %s' % (d['value'], d['ms_per_step'], r['frac'], r['traffic'], d['e2e']['value'], d['gpu_launches'], d['clocks']))
print('check', {k:v for k,v in c.items() if k!='rho_sys_final'})
print('cpu_baseline', d.get('cpu_baseline'))
print('other', d.get('other_workloads'))
PY
ls "$out"
