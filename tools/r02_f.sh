#!/bin/bash
O=gpurun_out; mkdir -p $O
python tools/bench_vec.py C5 592
python tools/bench_dos.py C5 592 1 1 200 ldos
python bench.py --workload C5 --no-e2e --no-cpu --steps 1 --warmup 1 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C5 value %.4e' % d['value'], json.dumps(d['secondary'])[:500])"
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee $O/f_pytest.log
