#!/bin/bash
O=gpurun_out; mkdir -p $O; : > $O/ts21.log
timeout 600 python -m pytest tests/test_gpu_twostage.py -x -q 2>&1 | tail -2 | tee -a $O/ts21.log
TS_TIME=592 python tools/test_twostage.py 256,384,512 2 2>&1 | grep "stage1.*ms" | cut -c1-150 | tee -a $O/ts21.log
python tools/bench_heev.py 512 4096 2>&1 | tail -1 | cut -c1-110 | tee -a $O/ts21.log
