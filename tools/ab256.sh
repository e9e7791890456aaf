#!/bin/bash
for v in "$@"; do
  env $v python bench.py --no-e2e --no-cpu --mesh 256 256 256 --steps 2 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$v', '%.4g' % d['value'], '%.4f' % d['roofline']['frac'])"
done
