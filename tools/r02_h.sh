#!/bin/bash
O=gpurun_out; mkdir -p $O
python tools/debug_c5b.py 2>&1 | tail -3
python tools/debug_bigvec.py 96,257,512 2>&1 | tail -10
python tools/bench_dos.py C5 2048 1 1 200 ldos
python tools/bench_vec.py C5 592
python tools/bench_heev.py 128,256 296 vec
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/h_pytest.log
