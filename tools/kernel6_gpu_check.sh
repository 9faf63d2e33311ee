#!/bin/bash
# First GPU call for kernels 6 and 7: parity tests, bench lines for the
# default kernel and for kernels 6 / 7, launch lists and one full ncu capture of each.
#
#   gpurun --timeout 1500 -- 'bash tools/kernel6_gpu_check.sh'
#
# Everything lands under gpurun_out/k6/; copy what is worth keeping into profiles/.
#
# Two GPUs (fused bulk-store halo push of kernel 6, and kernel 6 as the stage kernel of a sharded run):
#   gpurun --gpus 2 --timeout 900 -- 'PYQED_B200_TEST_KERNEL6=1 python -m pytest tests/test_gpu_kernel6.py -m gpu -q -k "fused or sharded";
#     for f in 0 1; do python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
#       bench.py --gpus 2 --kernel 6 --fused $f --steps 10 --warmup 3 > gpurun_out/k6/bench_n2_kernel6_fused$f.json; done'

set -u
out=gpurun_out/k6
mkdir -p "$out"

echo "== kernel 6 parity tests"
PYQED_B200_TEST_KERNEL6=1 timeout 600 python -m pytest tests/test_gpu_kernel6.py -m gpu -x -q 2>&1 | tail -15 | tee "$out/tests.log"

echo "== bench: default kernel, then kernels 6 and 7 (no CPU legs)"
PYQED_B200_BENCH_KERNEL6=0 timeout 600 python bench.py --no-cpu --steps 20 --warmup 3 > "$out/bench_default.json" 2> "$out/bench_default.err"
timeout 600 python bench.py --no-cpu --kernel 6 --steps 20 --warmup 3 > "$out/bench_kernel6.json" 2> "$out/bench_kernel6.err"
tail -c 1200 "$out/bench_default.json"; echo
tail -c 1200 "$out/bench_kernel6.json"; echo
timeout 600 python bench.py --no-cpu --kernel 7 --steps 20 --warmup 3 > "$out/bench_kernel7.json" 2> "$out/bench_kernel7.err"
tail -c 1200 "$out/bench_kernel7.json"; echo
timeout 600 python bench.py --no-cpu --kernel 7 --prefetch 1 --steps 20 --warmup 3 > "$out/bench_kernel7_prefetch.json" 2> "$out/bench_kernel7_prefetch.err"
tail -c 1200 "$out/bench_kernel7_prefetch.json"; echo

echo "== launch list of kernel 6 (per-launch times are cold-cache and serialised)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file "$out/launches_kernel6.csv" python bench.py --no-cpu --kernel 6 --steps 2 --warmup 3 \
    > "$out/ncu_launches.log" 2>&1

echo "== full capture of three launches of kernel 6 (middle, middle, last stage)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stage_rows_sym_kernel \
    -s 13 -c 3 -o "$out/kernel6_full" python bench.py --no-cpu --kernel 6 --steps 2 --warmup 3 \
    > "$out/ncu_full.log" 2>&1
echo "== the same for kernel 7 (packed storage)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file "$out/launches_kernel7.csv" python bench.py --no-cpu --kernel 7 --steps 2 --warmup 3 \
    > "$out/ncu_launches7.log" 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stage_rows_sym_kernel \
    -s 13 -c 3 -o "$out/kernel7_full" python bench.py --no-cpu --kernel 7 --steps 2 --warmup 3 \
    > "$out/ncu_full7.log" 2>&1
ls -la "$out"
