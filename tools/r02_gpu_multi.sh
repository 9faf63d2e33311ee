#!/bin/bash
# usage: tools/r02_gpu_multi.sh <outdir> <ngpus> [steps]   (run under gpurun --gpus N)
set -u
out=gpurun_out/${1:-multi}
N=${2:-2}
K=${3:-10}
mkdir -p "$out"
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > "$out/smi.txt"
nvidia-smi topo -m > "$out/topo.txt" 2>&1
echo "== 2-GPU tests (NCCL, device barrier)"
timeout 900 python -m pytest tests/test_sharded.py tests/test_gpu_kernel6.py -m gpu -q -k "two_gpus" 2>&1 | tail -6 | tee "$out/tests.log"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1]); r=d['roofline']; c=d['check']
    print(sys.argv[2], 'N=%d'%d['n_gpus'], r['kernel'], '%.4g ADO-steps/s' % d['value'], '%.3f ms/step' % d['ms_per_step'], 'stage %.3f ms' % r['avg_launch_ms'], 'frac %.3f agg %.3f' % (r['frac'] or 0, r['whole_job_frac_of_aggregate_peak']), 'e2e %.4g' % d['e2e']['value'], 'setup %.2f' % d['config']['setup_s_first_call'], 'vs_n1', c.get('max_abs_diff_vs_n1'), 'fixture', c.get('reference_fixture'))
    if 'ranks' in d: print('   ranks', d['ranks'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e, open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
}
echo "== 1 GPU (reference values of rho_sys for the 1-vs-N check)"
for k in 10 20; do
  timeout 300 python bench.py --no-cpu --steps $k --warmup 3 --save-rho-ref "$out/rho_sys_n1.json" > "$out/bench_n1_k$k.json" 2> "$out/bench_n1_k$k.err"; show "$out/bench_n1_k$k.json" n1_k$k
done
mkdir -p profiles; cp "$out/rho_sys_n1.json" profiles/r02_rho_sys_n1.json
tr() { label=$1; n=$2; shift; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --no-cpu --steps $K --warmup 3 "$@" > "$out/bench_$label.json" 2> "$out/bench_$label.err"; show "$out/bench_$label.json" $label; }
n=2
while [ $n -le $N ]; do
  tr native_n$n $n
  n=$((n*2))
done
tr legacy_n$N $N --native 0
tr native_k6_n$N $N --kernel 6
ls "$out"
