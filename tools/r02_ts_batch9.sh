#!/bin/bash
O=gpurun_out; mkdir -p $O; : > $O/ts19.log
python tools/bench_heev.py 224,256,320,384,448,512 592 2>&1 | grep "N=" | cut -c1-110 | tee -a $O/ts19.log
python tools/bench_heev.py 512 4096 2>&1 | tail -1 | cut -c1-110 | tee -a $O/ts19.log
python tools/bench_dos.py C5 2048 1 1 200 ldos 2>&1 | tail -1 | cut -c1-110 | tee -a $O/ts19.log
python bench.py --workload C5 --no-cpu --steps 3 --warmup 3 2>/dev/null | tail -1 > $O/ts19_bench_C5.json
python -c "
import json; d=json.load(open('$O/ts19_bench_C5.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['secondary']['edge_ldos']['value'])" | tee -a $O/ts19.log
timeout 600 python -m pytest tests/test_gpu_twostage.py -x -q 2>&1 | tail -2 | tee -a $O/ts19.log
