#!/bin/bash
set -u
out=gpurun_out/${1:-n2check}; mkdir -p "$out"
timeout 900 python -m pytest tests/test_sharded.py tests/test_gpu_kernel6.py -m gpu -q -k "two_gpus" 2>&1 | tail -4 | tee "$out/tests.log"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1]); r=d['roofline']; c=d['check']
    print(sys.argv[2], 'N=%d'%d['n_gpus'], r['kernel'], '%.4g ADO-steps/s' % d['value'], '%.3f ms/step' % d['ms_per_step'], 'stage %.3f ms' % r['avg_launch_ms'], 'agg %.3f' % r['whole_job_frac_of_aggregate_peak'], 'setup %.2f' % d['config']['setup_s_first_call'], 'vs_n1', c.get('max_abs_diff_vs_n1'), 'fixture ok', c.get('reference_fixture',{}).get('ok'), 'rebalance', (d['config'].get('setup_breakdown_s') or {}).get('rebalance',{}).get('changed'))
except Exception as e:
    print(sys.argv[2], 'FAILED', e, open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
}
tr() { label=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --no-cpu --warmup 3 "$@" > "$out/bench_$label.json" 2> "$out/bench_$label.err"; show "$out/bench_$label.json" $label; }
tr n2 --steps 20

timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --no-cpu --warmup 3 --workload aggregate7_K6_L6 --batch 64 --steps 700 > "$out/bench_n2_batch.json" 2> "$out/bench_n2_batch.err"; tail -c 600 "$out/bench_n2_batch.json"
