#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee $O/c_pytest.log
{
python tools/bench_heev.py 96,128,256 296 vec
python tools/bench_heev.py 512 148 vec
SKTB_BIGVEC=0 python tools/bench_heev.py 128,256 296 vec
python tools/bench_vec.py C5 592
python tools/bench_dos.py C5 592 1 1 200 ldos
python bench.py --workload C5 --no-e2e --no-cpu --steps 1 --warmup 1 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C5 value %.4e' % d['value'], json.dumps(d['secondary'])[:600])"
} 2>&1 | tee $O/c_bench.log
