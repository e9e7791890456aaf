#!/bin/bash
O=gpurun_out; mkdir -p $O
{ for r in 1 2; do python bench.py --workload C4 --no-cpu --no-e2e --no-secondary --steps 2 --warmup 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('C4', '%.5g' % d['value'], '%.4f' % d['roofline']['frac'])"; done
  python tools/bench_heev.py 96,128,160,192 2048; python tools/bench_heev.py 48,64 8192
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_knobs.py -m gpu -q -x 2>&1 | tail -3; } > $O/s3_ab2.log 2>&1
cat $O/s3_ab2.log
