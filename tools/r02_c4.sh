#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/c4_pytest.log
for cfg in default SKTB_CTA64_FUSE=0; do
  envs=$cfg; [ "$cfg" = "default" ] && envs="X=1"
  echo "== $cfg"
  env $envs python bench.py --workload C4 --mesh 96 96 96 --steps 2 --warmup 2 --no-cpu --no-e2e --no-secondary 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4e k/s  frac %.4f  ms %.2f  %s' % (d['value'], d['roofline']['frac'], d['ms_per_step'], d['roofline']['kernel'][:60]))"
  env $envs python tools/bench_dos.py C4 28 28 28 400 ldos 2>&1 | tail -2
done 2>&1 | tee $O/c4_ab.log
