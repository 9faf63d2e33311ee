#!/bin/bash
# usage: tools/r02_gpu_shard_probe.sh <outdir> <world> <profiled rank>
set -u
out=gpurun_out/${1:-probe}; N=${2:-8}; P=${3:-3}
mkdir -p "$out"
export WORLD_SIZE=$N MASTER_ADDR=127.0.0.1 MASTER_PORT=29611 LOCAL_RANK=0
pids=""
for r in $(seq 0 $((N-1))); do
  if [ $r -eq $P ]; then
    RANK=$r timeout 900 ncu --set full --clock-control none --import-source on -k regex:stage_rows_sym_kernel -s 5 -c 2 -o "$out/shard_rank$P" \
        python tools/shard_probe.py > "$out/rank$r.log" 2>&1 &
  else
    RANK=$r timeout 900 python tools/shard_probe.py > "$out/rank$r.log" 2>&1 &
  fi
  pids="$pids $!"
done
for p in $pids; do wait $p; done
tail -n 2 "$out"/rank*.log
ncu -i "$out/shard_rank$P.ncu-rep" --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
for d in rows[2:]:
    print('----', d[hdr.index('Kernel Name')][:80])
    for w in ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','launch__registers_per_thread','launch__block_size','launch__shared_mem_per_block_dynamic']:
        if w in hdr: print('  ', w, d[hdr.index(w)])
    st=sorted(((float(d[hdr.index(h)].replace(',','')) if d[hdr.index(h)] else 0,h) for h in hdr if 'average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio')),reverse=True)
    for v,h in st[:6]: print('     %.3f %s'%(v,h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
"
