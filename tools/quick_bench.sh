#!/bin/bash
# usage: tools/quick_bench.sh <label> [bench args...]   (prints label, value, ms/step, roofline frac)
label=$1; shift
python bench.py --no-cpu "$@" 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('$label', r['kernel'], '%.4g ADO-steps/s' % d['value'], '%.3f ms/step' % d['ms_per_step'], 'frac %.3f' % (r['frac'] or 0), 'e2e %.4g' % d['e2e']['value'])
except Exception as e:
    print('$label', 'FAILED', e)
"
