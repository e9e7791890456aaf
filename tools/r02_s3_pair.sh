#!/bin/bash
O=gpurun_out; mkdir -p $O
{ SKTB_GF_PAIR=1 timeout 600 python -m pytest tests/test_gpu_gf_inverse.py -m gpu -q 2>&1 | tail -6
  for r in 1 2; do python tools/bench_dos.py T 20 20 20 400 dos auto; SKTB_GF_PAIR=1 python tools/bench_dos.py T 20 20 20 400 dos auto; done; } > $O/s3_pair.log 2>&1
cat $O/s3_pair.log | cut -c1-200
