#!/bin/bash
O=gpurun_out; mkdir -p $O
bash tools/ab_lib.sh pysktb_b200/libsktb_base.so base > $O/s3_ab.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_knobs.py -m gpu -q -x 2>&1 | tail -3 >> $O/s3_ab.log
python bench.py --workload C4 --no-cpu --no-e2e --no-secondary --steps 2 --warmup 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('C4', '%.5g' % d['value'])" >> $O/s3_ab.log
cat $O/s3_ab.log
