#!/bin/bash
# session 3: dense-inverse Green's function - new GPU tests, the touched parity test, timing of the inverse path (tile kernels and, with SKTB_GF_TILE=0, the warp / CTA kernels)
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gf_inverse.py -m gpu -q 2>&1 | tail -25 > $O/s3_gf_pytest.log
SKTB_GF_TILE=0 timeout 900 python -m pytest tests/test_gpu_gf_inverse.py -m gpu -q 2>&1 | tail -5 >> $O/s3_gf_pytest.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "nonherm or dos or ldos or retarded or spectral" 2>&1 | tail -8 >> $O/s3_gf_pytest.log
{ python tools/bench_dos.py T 20 20 20 400 dos inverse; python tools/bench_dos.py C4 10 10 10 400 dos inverse; python tools/bench_dos.py T 20 20 20 400 ldos inverse;
  SKTB_GF_TILE=0 python tools/bench_dos.py T 20 20 20 400 dos inverse; SKTB_GF_TILE=0 python tools/bench_dos.py C4 10 10 10 400 dos inverse; } > $O/s3_gf_bench.log 2>&1
cat $O/s3_gf_pytest.log $O/s3_gf_bench.log
