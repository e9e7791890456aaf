#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_twostage.py -x -q 2>&1 | tail -15 | tee $O/ts12_pytest.log
python bench.py --workload C5 --steps 3 --warmup 3 --no-cpu 2>$O/ts12_bench_C5.err | tail -1 > $O/ts12_bench_C5.json
python -c "
import json; d=json.load(open('$O/ts12_bench_C5.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac']); print(d['secondary']['edge_ldos']['value'], d['secondary']['edge_ldos']['kernel'][:200])" | tee $O/ts12_c5.log
for n in 888 4096; do python tools/bench_heev.py 512 $n 2>&1 | tail -1; done | tee $O/ts12_big.log
