#!/bin/bash
O=gpurun_out; mkdir -p $O
for rep in 1 2; do for v in "SKTB_PAIR_CHUNK_LOG2=20" "SKTB_PAIR_CHUNK_LOG2=19" "SKTB_PAIR_CHUNK_LOG2=21" "SKTB_PAIR_CHUNK_LOG2=18"; do
  env $v python bench.py --no-e2e --no-cpu --no-secondary --mesh 256 256 256 --steps 3 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$v', '%.5g' % d['value'], '%.4f' % d['roofline']['frac'])"
done; done > $O/s3_env.log 2>&1
cat $O/s3_env.log
