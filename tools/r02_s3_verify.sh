#!/bin/bash
# session-3 check of HEAD: GPU tests, smoke, C5 bench line (the last commit touched stage 1)
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/s3_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/s3_smoke.log 2>&1
python bench.py --workload C5 --no-cpu --steps 3 --warmup 3 2>/dev/null | tail -1 > $O/s3_bench_C5.json
cat $O/s3_pytest.log; tail -3 $O/s3_smoke.log; python -c "
import json; d=json.load(open('$O/s3_bench_C5.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['secondary']['edge_ldos']['value'])"
