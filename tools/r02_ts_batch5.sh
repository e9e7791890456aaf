#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_twostage.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4 | tee $O/ts15_pytest.log
python tools/bench_heev.py 128,256 296 vec 2>&1 | tail -2 | cut -c1-100 | tee $O/ts15_vec.log
python tools/bench_heev.py 512 148 vec 2>&1 | tail -1 | cut -c1-100 | tee -a $O/ts15_vec.log
python tools/bench_dos.py C5 2048 1 1 200 ldos 2>&1 | tail -1 | cut -c1-120 | tee $O/ts15_ldos.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/ts15_launches.csv python tools/bench_dos.py C5 888 1 1 200 ldos > /dev/null 2>&1
python tools/launch_breakdown.py $O/ts15_launches.csv | grep -E "stein|stage|rows" | cut -c1-100 | tee $O/ts15_breakdown.log
