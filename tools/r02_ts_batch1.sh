#!/bin/bash
O=gpurun_out; mkdir -p $O
TS_TIME=296 timeout 200 python tools/test_twostage.py 128,192,256,320,384,512 2 2>&1 | grep stage1.*ms | tee $O/ts11_time.log
SKTB_TWOSTAGE_MIN=9999 python tools/bench_heev.py 128,192,256,320 296 2>&1 | tail -4 | tee $O/ts11_old.log
for n in 1184 4096; do python tools/bench_heev.py 512 $n 2>&1 | tail -1; done | tee $O/ts11_big.log
timeout 900 python -m pytest tests/test_gpu_twostage.py tests/test_gpu_knobs.py -x -q 2>&1 | tail -4 | tee $O/ts11_pytest.log
