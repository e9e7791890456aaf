#!/bin/bash
O=gpurun_out; mkdir -p $O
echo "== default (3 CTAs/SM, 80 regs)" | tee $O/ts16.log
TS_TIME=444 python tools/test_twostage.py 512 2 2>&1 | grep "stage1.*ms" | cut -c1-150 | tee -a $O/ts16.log
TS_TIME=592 python tools/test_twostage.py 512 2 2>&1 | grep "stage1.*ms" | cut -c1-150 | tee -a $O/ts16.log
TS_TIME=888 python tools/test_twostage.py 512 2 2>&1 | grep "stage1.*ms" | cut -c1-150 | tee -a $O/ts16.log
touch pysktb_b200/csrc/heev_twostage.cu; make -C pysktb_b200/csrc TSEXTRA=-DS2_MIN_CTAS=4 2>&1 | grep -E "error" 
echo "== 4 CTAs/SM, 64 regs" | tee -a $O/ts16.log
TS_TIME=592 python tools/test_twostage.py 512 2 2>&1 | grep "stage1.*ms" | cut -c1-150 | tee -a $O/ts16.log
TS_TIME=1184 python tools/test_twostage.py 512 2 2>&1 | grep "stage1.*ms" | cut -c1-150 | tee -a $O/ts16.log
touch pysktb_b200/csrc/heev_twostage.cu; make -C pysktb_b200/csrc TSEXTRA=-DS2_MIN_CTAS=2 2>&1 | grep -E "error"
echo "== 2 CTAs/SM, up to 128 regs" | tee -a $O/ts16.log
TS_TIME=592 python tools/test_twostage.py 512 2 2>&1 | grep "stage1.*ms" | cut -c1-150 | tee -a $O/ts16.log
