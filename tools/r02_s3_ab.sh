#!/bin/bash
O=gpurun_out; mkdir -p $O
bash tools/ab_lib.sh base pysktb_b200/libsktb_si.so > $O/s3_ab.log 2>&1
for l in "" "$PWD/pysktb_b200/libsktb_si.so"; do for w in C1 C3 C4; do SKTB_LIB_PATH=$l python bench.py --workload $w --no-cpu --no-e2e --no-secondary --steps 2 --warmup 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$w', '${l: -14}', '%.5g' % d['value'])"; done; done >> $O/s3_ab.log 2>&1
SKTB_LIB_PATH=$PWD/pysktb_b200/libsktb_si.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3 >> $O/s3_ab.log
cat $O/s3_ab.log
