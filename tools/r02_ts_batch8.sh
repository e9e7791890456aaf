#!/bin/bash
O=gpurun_out; mkdir -p $O; : > $O/ts18.log
for cfg in "4 4" "2 8" "4 3"; do
  set -- $cfg
  touch pysktb_b200/csrc/heev_twostage.cu; make -C pysktb_b200/csrc TSEXTRA="-DS2_MIN_CTAS=$2 -DS2_NWARPS=$1" 2>&1 | grep -E "error"
  echo "== $2 CTAs/SM x $1 warps" | tee -a $O/ts18.log
  for n in 592 1184; do TS_TIME=$n python tools/test_twostage.py 512 2 2>&1 | grep "stage1.*ms" | cut -c1-150 | tee -a $O/ts18.log; done
  TS_TIME=1184 python tools/test_twostage.py 256 2 2>&1 | grep "stage1.*ms" | cut -c1-150 | tee -a $O/ts18.log
done
