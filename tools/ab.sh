#!/bin/bash
# A/B runs of bench.py (device-resident value only) over environment variants.  usage: tools/ab.sh "VAR=a VAR2=b" "VAR=c" ...
for v in "$@"; do
  env $v python bench.py --no-e2e --no-cpu --mesh 128 128 128 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$v', '%.4g' % d['value'], '%.4f' % d['roofline']['frac'], d['roofline']['kernel'][:60])"
done
