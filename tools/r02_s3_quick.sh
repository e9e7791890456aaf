#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gf_inverse.py -m gpu -q 2>&1 | tail -12 > $O/s3_q_pytest.log
python bench.py --workload C4 --no-cpu --steps 1 --warmup 1 2>/dev/null | tail -1 > $O/s3_q_bench_C4.json
cat $O/s3_q_pytest.log; python -c "
import json; d=json.load(open('$O/s3_q_bench_C4.json')); print(d['value'], {k:(v.get('value'), v.get('roofline',{}).get('frac')) for k,v in d['secondary'].items()})"
