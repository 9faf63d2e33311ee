#!/bin/bash
# Round 2 evidence run (one GPU): launch list + full capture of the shipped stage kernel, compute-sanitizer
set -u
out=gpurun_out/${1:-r02evidence}
mkdir -p "$out"
echo "== launch list of the default bench command (cold-cache, serialised launches)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches.csv" \
    python bench.py --no-cpu --steps 2 --warmup 3 > "$out/launches.log" 2>&1
python - "$out/launches.csv" <<'PY'
import csv,sys,collections
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0][:70]; t=float(r[-1])
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=t
tot=sum(a[1] for a in agg.values())
for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1]): print('%6d launches %12.1f us %5.1f%%  %s'%(n,t/1e3,100*t/tot,k))
PY
echo "== full capture: kernel 7 middle and last stage"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage_rows_sym_kernel -s 14 -c 2 -o "$out/k7_final" \
    python bench.py --no-cpu --steps 2 --warmup 3 > "$out/ncu_full.log" 2>&1
echo "== compute-sanitizer"
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernel6.py -m gpu -q -x \
     -k "test_deom_matches_reference and (fmo_K21_L2 or random5 or spin_boson or aggregate_L3_T0 or polariton8) or test_kernel6_matches_reference and K21_L2 and 2" 2>&1 | grep -E "ERROR SUMMARY|passed|failed" | tee -a "$out/sanitizer_$tool.txt"
done
timeout 900 compute-sanitizer --tool racecheck python tests/tools/sanitizer_case.py 2>&1 | grep -E "RACECHECK SUMMARY|^ok|Error|hazard" | tee "$out/sanitizer_racecheck.txt"
ls -la "$out"
