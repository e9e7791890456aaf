"""K1 alone (sktb_hk, materialising H(k)): achieved HBM GB/s against the algorithmic 16 N^2 + 24 bytes per k-point.
usage: bench_hk.py [workload ...]   (default: C4 C5 T)"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pysktb_b200 as tb
from pysktb_b200 import workloads as W, _lib
from bench import WORKLOADS

peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
hbm_peak = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
lib = _lib.lib()
for name in (sys.argv[1:] or ["C4", "C5", "T"]):
    builder, kw, _ = WORKLOADS[name]
    spec = W.ALL[builder](**kw)
    _, ham = W.instantiate(spec, tb, tb.scaling)
    N = len([o for _, _, orbs in spec["atoms"] for o in orbs]) * (2 if spec["soc"] else 1)
    nk = int(min(1 << 18, (6 << 30) // (16 * N * N)))
    k = torch.from_numpy(np.random.default_rng(0).random((nk, 3))).cuda()
    H = torch.empty((nk, N, N), dtype=torch.complex128, device="cuda")
    run = lambda: _lib.check(lib.sktb_hk(ham._plan, k.data_ptr(), nk, int(spec["soc"]), H.data_ptr(), None, None))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gbs = nk * (16.0 * N * N + 24) / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": "hk_kernel", "workload": name, "N": N, "k_points": nk, "ms": ms, "k_per_s": nk / (ms * 1e-3),
                      "algorithmic_bytes_per_k": 16 * N * N + 24, "achieved_gbs": gbs, "hbm_peak_gbs": hbm_peak, "frac": gbs / hbm_peak,
                      "note": "output %.2f GB per launch >> 126 MB L2" % (nk * 16.0 * N * N / 1e9)}))
