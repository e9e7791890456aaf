#!/bin/bash
O=gpurun_out; mkdir -p $O
python tools/debug_c5b.py 2>&1 | tail -8
python tools/bench_heev.py 256 296 vec
python tools/bench_heev.py 512 148 vec
python tools/bench_vec.py C5 592
python tools/bench_dos.py C5 592 1 1 200 ldos
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee $O/e_pytest.log
