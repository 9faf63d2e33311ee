#!/bin/bash
# Round 2, second final check on one GPU (after kernel 9): what the driver runs (GPU tests, smoke, default bench),
# config 4 / config 5 bench lines, ncu captures of kernel 9 and of the batched resident kernel, compute-sanitizer
set -u
out=gpurun_out/${1:-r02final2}
mkdir -p "$out"
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -rfs 2>&1 | tail -8 | tee "$out/tests.log"
echo "== smoke"
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -9 | tee "$out/smoke.log"
b() { label=$1; shift; timeout 300 python bench.py --no-cpu --warmup 3 "$@" > "$out/bench_$label.json" 2> "$out/bench_$label.err"; python - "$out/bench_$label.json" "$label" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1]); r=d['roofline']
    print(sys.argv[2], d['config']['workload'], 'batch', d['config'].get('batch',1), r['kernel'], '%.4g ADO-steps/s' % d['value'], '%.2f us/step' % (1e3*d['ms_per_step']), 'e2e %.4g' % d['e2e']['value'], 'frac', r.get('frac'))
except Exception as e:
    print(sys.argv[2], 'FAILED', e, open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
}
echo "== small configurations"
b cfg4_k9 --workload polariton32_K4_L6 --steps 3000
b cfg4_k8 --workload polariton32_K4_L6 --steps 3000 --kernel 8
b cfg4_k2 --workload polariton32_K4_L6 --steps 1000 --kernel 2
b cfg5_b64 --workload aggregate7_K6_L6 --batch 64 --steps 700
b cfg1 --workload spin_boson_K2_L10 --steps 3000
b cfg2 --workload fmo7_K7_L4 --steps 3000
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
echo "== ncu kernel 9"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:stage_dataflow_tma -s 1 -c 1 -f -o "$out/k9_full" python bench.py --no-cpu --workload polariton32_K4_L6 --steps 300 --warmup 1 > "$out/ncu_k9.log" 2>&1
tail -2 "$out/ncu_k9.log"
echo "== ncu resident kernel, config 5 batch 64"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:resident_elem -s 1 -c 1 -f -o "$out/cfg5_full" python bench.py --no-cpu --workload aggregate7_K6_L6 --batch 64 --steps 200 --warmup 1 > "$out/ncu_cfg5.log" 2>&1
tail -2 "$out/ncu_cfg5.log"
echo "== compute-sanitizer (memcheck, racecheck) on the tiny case incl. kernels 8 and 9"
export PYQED_HEOM_DATAFLOW_TIMEOUT_MS=600000
timeout 600 compute-sanitizer --tool memcheck python tests/tools/sanitizer_case.py 2>&1 | grep -E "ERROR SUMMARY|^ok|Error|error" | tee "$out/sanitizer_memcheck.txt"
timeout 900 compute-sanitizer --tool racecheck python tests/tools/sanitizer_case.py 2>&1 | grep -E "RACECHECK SUMMARY|^ok|Error|hazard" | tee "$out/sanitizer_racecheck.txt"
ls "$out"
