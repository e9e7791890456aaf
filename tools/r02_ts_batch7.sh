#!/bin/bash
O=gpurun_out; mkdir -p $O
echo "== default (2 CTAs/SM x 8 warps)" | tee $O/ts17.log
for n in 296 592 888; do TS_TIME=$n python tools/test_twostage.py 512 2 2>&1 | grep "stage1.*ms" | cut -c1-150 | tee -a $O/ts17.log; done
TS_TIME=592 python tools/test_twostage.py 256,384 2 2>&1 | grep "stage1.*ms" | cut -c1-150 | tee -a $O/ts17.log
python tools/bench_heev.py 512 4096 2>&1 | tail -1 | cut -c1-110 | tee -a $O/ts17.log
python tools/bench_dos.py C5 2048 1 1 200 ldos 2>&1 | tail -1 | cut -c1-110 | tee -a $O/ts17.log
touch pysktb_b200/csrc/heev_twostage.cu; make -C pysktb_b200/csrc TSEXTRA="-DS2_MIN_CTAS=1 -DS2_NWARPS=16" 2>&1 | grep -E "error"
echo "== 1 CTA/SM x 16 warps" | tee -a $O/ts17.log
for n in 148 296 592; do TS_TIME=$n python tools/test_twostage.py 512 2 2>&1 | grep "stage1.*ms" | cut -c1-150 | tee -a $O/ts17.log; done
TS_TIME=592 python tools/test_twostage.py 256,384 2 2>&1 | grep "stage1.*ms" | cut -c1-150 | tee -a $O/ts17.log
