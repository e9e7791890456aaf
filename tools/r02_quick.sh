#!/bin/bash
# quick A/B on the headline path: parity subset + T at 256^3 with the knobs given as arguments ("NAME=VAL,NAME=VAL" per run)
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/quick_pytest.log
cat $O/quick_pytest.log
for cfg in "$@"; do
  envs=$(echo "$cfg" | tr ',' ' ')
  [ "$cfg" = "default" ] && envs=""
  for rep in 1 2; do
    echo "== $cfg rep $rep"
    env $envs python bench.py --mesh 256 256 256 --steps 3 --warmup 3 --no-cpu --no-e2e --no-secondary 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4e k/s  frac %.4f  ms %.2f  qlfail %s' % (d['value'], d['roofline']['frac'], d['ms_per_step'], d['ql_failures']))"
  done
done 2>&1 | tee $O/quick_ab.log
