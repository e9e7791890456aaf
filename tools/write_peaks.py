"""Measure the FP64 (DFMA) peak the roofline fractions are quoted against and record it with its method.
usage: write_peaks.py out.json"""
import sys, os, json, datetime
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pysktb_b200 import _lib
lib = _lib.lib()
vals = []
for _ in range(5):
    t, ms = _lib.C.c_double(), _lib.C.c_double()
    _lib.check(lib.sktb_fp64_peak(0, _lib.C.byref(t), _lib.C.byref(ms)))
    vals.append(t.value)
p = torch.cuda.get_device_properties(0)
out = {"fp64_tflops": max(vals), "samples_tflops": vals, "gpu_name": p.name, "sms": p.multi_processor_count,
       "nominal_tflops": p.multi_processor_count * 64 * 2 * 1.965e9 / 1e12,
       "method": "sktb_fp64_peak (pysktb_b200/csrc/sktb.cu, fp64_peak_kernel in kernels_generic.cuh): 4 CTAs x 512 threads per SM, every thread runs "
                 "65536 iterations of 8 independent DFMA chains (2 flops each, operands reused from registers), CUDA events around the launch, best of 3 "
                 "launches after a warm-up; this file: best of 5 calls.  bench.py measures the same figure live on the benched device for roofline.peak.",
       "when": datetime.datetime.utcnow().isoformat() + "Z"}
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps(out))
