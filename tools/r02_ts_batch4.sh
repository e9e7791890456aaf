#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_twostage.py tests/test_gpu_parity.py -x -q 2>&1 | tail -6 | tee $O/ts14_pytest.log
python bench.py --workload C5 --steps 3 --warmup 3 --no-cpu 2>$O/ts14_bench_C5.err | tail -1 > $O/ts14_bench_C5.json
python -c "
import json; d=json.load(open('$O/ts14_bench_C5.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac']); print(d['secondary']['edge_ldos']['value'], d['secondary']['edge_ldos']['kernel'][:120])" | tee $O/ts14_c5.log
python tools/bench_heev.py 256,512 444 vec 2>&1 | tail -2 | tee $O/ts14_vec.log
