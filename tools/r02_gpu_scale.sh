#!/bin/bash
# usage: tools/r02_gpu_scale.sh <outdir> <ngpus> [steps]   scaling run: native sharded bench at N = 2, 4, .., ngpus
set -u
out=gpurun_out/${1:-scale}
N=${2:-8}
K=${3:-10}
mkdir -p "$out"
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > "$out/smi.txt"
nvidia-smi topo -m > "$out/topo.txt" 2>&1
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1]); r=d['roofline']; c=d['check']
    print(sys.argv[2], 'N=%d'%d['n_gpus'], r['kernel'], '%.4g ADO-steps/s' % d['value'], '%.3f ms/step' % d['ms_per_step'], 'stage %.3f ms' % r['avg_launch_ms'], 'frac %.3f agg %.3f' % (r['frac'] or 0, r['whole_job_frac_of_aggregate_peak']), 'e2e %.4g' % d['e2e']['value'], 'setup %.2f' % d['config']['setup_s_first_call'], 'vs_n1', c.get('max_abs_diff_vs_n1'), 'fixture ok', c.get('reference_fixture',{}).get('ok'))
    print('   setup', d['config'].get('setup_breakdown_s'))
    if 'ranks' in d: print('   ranks', [(x['owned_ados'], x['halo_bytes_per_stage']//1000000, round(x['avg_stage_kernel_ms'],3)) for x in d['ranks']])
except Exception as e:
    print(sys.argv[2], 'FAILED', e, open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
}
tr() { label=$1; n=$2; shift; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --no-cpu --warmup 3 "$@" > "$out/bench_$label.json" 2> "$out/bench_$label.err"; show "$out/bench_$label.json" $label; }
timeout 300 python bench.py --no-cpu --steps $K --warmup 3 > "$out/bench_n1.json" 2> "$out/bench_n1.err"; show "$out/bench_n1.json" n1
n=2
while [ $n -le $N ]; do
  tr native_n$n $n --steps $K
  n=$((n*2))
done
tr native_n${N}_k20 $N --steps 20
ls "$out"
