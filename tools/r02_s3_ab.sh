#!/bin/bash
O=gpurun_out; mkdir -p $O
bash tools/ab_lib.sh base pysktb_b200/libsktb_g2.so > $O/s3_ab.log 2>&1
SKTB_LIB_PATH=$PWD/pysktb_b200/libsktb_g2.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3 >> $O/s3_ab.log
cat $O/s3_ab.log
