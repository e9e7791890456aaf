#!/bin/bash
O=gpurun_out; mkdir -p $O; : > $O/ts20.log
SKTB_TWOSTAGE_MIN=96 python tools/bench_heev.py 128,160,176,192,208 592 2>&1 | grep "N=" | cut -c1-100 | tee -a $O/ts20.log
SKTB_TWOSTAGE_MIN=100000 python tools/bench_heev.py 128,160,176,192,208,224 592 2>&1 | grep "N=" | cut -c1-100 | tee -a $O/ts20.log
SKTB_TWOSTAGE_MIN=96 python tools/bench_heev.py 128,160,192 8192 2>&1 | grep "N=" | cut -c1-100 | tee -a $O/ts20.log
SKTB_TWOSTAGE_MIN=100000 python tools/bench_heev.py 128,160,192 8192 2>&1 | grep "N=" | cut -c1-100 | tee -a $O/ts20.log
