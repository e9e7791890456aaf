#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/b_pytest.log
one() { # workload mesh... envs
  w=$1; shift; m="$1 $2 $3"; shift 3
  env "$@" python bench.py --workload $w --mesh $m --steps 2 --warmup 3 --no-cpu --no-e2e --no-secondary 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w $*: %.4e k/s  frac %.4f  ms %.2f  %s' % (d['value'], d['roofline']['frac'], d['ms_per_step'], d['roofline']['kernel'][:70]))"
}
{
one C4 96 96 96 X=1
one C4 96 96 96 SKTB_CTA64_FUSE=0
env X=1 python tools/bench_dos.py C4 28 28 28 400 ldos 2>&1 | tail -1
env SKTB_CTA64_FUSE=0 python tools/bench_dos.py C4 28 28 28 400 ldos 2>&1 | tail -1
one C1 256 256 256 X=1
one C1 256 256 256 SKTB_NO_THREAD=1
one T 256 256 256 X=1
one C3 1024 1024 1 X=1
} 2>&1 | tee $O/b_ab.log
