#!/bin/bash
# A/B of library builds with the same ABI (SKTB_LIB_PATH): device-resident T value on a 256^3 mesh, interleaved repetitions.
# usage: tools/ab_lib.sh lib1.so lib2.so ...   (paths relative to the repo root; "base" = pysktb_b200/libsktb.so)
for rep in 1 2; do
  for l in "$@"; do
    if [ "$l" = "base" ]; then p=""; else p="$PWD/$l"; fi
    SKTB_LIB_PATH=$p python bench.py --no-e2e --no-cpu --no-secondary --mesh 256 256 256 --steps 3 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$l', '%.5g' % d['value'], '%.4f' % d['roofline']['frac'])"
  done
done
