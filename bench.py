#!/usr/bin/env python
"""bench.py - k-points/sec of the H(k)+eigh hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload T|C1|C2|C3|C4|C5] [--impl reference]

One "step" = one pass of the fused K0+K1+K2 path over this rank's slice of a synthetic uniform k-mesh
(GreensFunction._generate_kmesh order).  Default workload: the north-star target model T (one Ce-like atom,
s+p+d+f with SOC, N = 32) on a 512^3 mesh per GPU (weak scaling: per-GPU work is fixed, the global mesh is
[512*n_gpus, 512, 512] split contiguously; no data-path collective).  Prints ONE JSON line (rank 0).

  value     device-resident throughput: k generated in-kernel, eigenvalues written to HBM.
  e2e       the same metric through the public API Hamiltonian.solve_kpath with page-locked HOST buffers:
            H2D of the k list and D2H of all eigenvalues inside the timed region (chunked, overlapped).
  roofline  FP64 CUDA-core bound of the eigensolver: (16/3) N^3 flops per k (SURVEY.md §8d convention) over the
            CUDA-event time of the kernel, against a DFMA microbenchmark measured live on the same device
            (MEASURED_PEAKS.json has no FP64 entry); the HBM side (8N bytes per k out) is reported beside it.
  cpu_baseline / --impl reference   the numpy oracle port of the reference's path (oracle/, the reference itself
            is pure Python and is not present on the GPU box) on the host cores, bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (workloads builder, kwargs, per-GPU mesh, eigenvectors?)
    "T": ("ce_spdf_1atom", {}, [512, 512, 512]),
    "C1": ("cubic_sp3", {}, [256, 256, 256]),
    "C2": ("graphene_pz", {}, [4096, 4096, 1]),
    "C3": ("sb_buckled", {}, [1024, 1024, 1]),
    "C4": ("ce_spdf_2atom", {}, [48, 48, 48]),
    "C5": ("zigzag_ribbon", {"n_chains": 256}, [592, 1, 1]),
}


def flops_per_k(N):
    return 16.0 / 3.0 * N ** 3


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max((float(r[3]) for r in self.rows if len(r) > 3 and r[3].replace(".", "", 1).isdigit()), default=None),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on all host cores
# ------------------------------------------------------------------------------------------------------
_ORC = None


def _cpu_init(spec_name, kw):
    global _ORC
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from pysktb_b200 import workloads as W
    from oracle.sktb_oracle import OracleModel
    from oracle import laws_oracle
    spec = W.ALL[spec_name](**kw)
    params = {k: {kk: (laws_oracle.make(v[1], **v[2]) if isinstance(v, tuple) and v and v[0] == "law" else v) for kk, v in t.items()}
              for k, t in spec["params"].items()}
    _ORC = (OracleModel(spec["lattice"], spec["a"], spec["atoms"], spec["bond_cut"], params, spec["periodicity"]), spec["soc"])


def _cpu_chunk(args):
    nk, lo, hi = args
    orc, soc = _ORC
    ks = [orc.kmesh_point(nk, i) for i in range(lo, hi)]
    return float(orc.solve_kpath(ks, soc=soc).sum())


def cpu_throughput(spec_name, kw, mesh, target_s=12.0, steps=1):
    """k-points/s of the oracle on all host cores over a bounded sample of the same mesh."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(spec_name, kw)) as pool:
        pool.map(_cpu_chunk, [(mesh, c * 4, c * 4 + 4) for c in range(cores)])       # warm-up (imports, model build)
        t0 = time.perf_counter()
        pool.map(_cpu_chunk, [(mesh, c * 16, c * 16 + 16) for c in range(cores)])    # calibration
        per_k = (time.perf_counter() - t0) / 16.0
        n = int(max(cores * 8, min(2_000_000, target_s / max(per_k, 1e-6) * cores)))
        n -= n % cores
        chunk = n // cores
        times = []
        for _ in range(steps):
            t0 = time.perf_counter()
            pool.map(_cpu_chunk, [(mesh, c * chunk, (c + 1) * chunk) for c in range(cores)])
            times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    return n / t, cores, f"first {n} points of the {mesh[0]}x{mesh[1]}x{mesh[2]} mesh, {cores} processes x {chunk} k", t


# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="T", choices=list(WORKLOADS))
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--mesh", type=int, nargs=3, default=None, help="override the per-GPU mesh")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    builder, kw, mesh = WORKLOADS[a.workload]
    mesh = list(a.mesh) if a.mesh else list(mesh)
    from pysktb_b200 import workloads as W
    spec = W.ALL[builder](**kw)
    N = len([o for _, _, orbs in spec["atoms"] for o in orbs]) * (2 if spec["soc"] else 1)
    per_gpu_k = mesh[0] * mesh[1] * mesh[2]
    config = {"workload": f"{a.workload}: {spec['name']}, N={N}, soc={spec['soc']}, eigenvalues only, "
                          f"uniform mesh {mesh[0]}x{mesh[1]}x{mesh[2]} per GPU ({per_gpu_k} k-points/GPU)",
              "k_points_per_gpu": per_gpu_k, "matrix_size": N, "parallelism": f"k-sharded x{a.gpus}",
              "cache": "no reusable inputs: k generated in-kernel, output (8N B/k) >> L2"}

    if a.impl == "reference":
        if rank != 0:
            return
        v, cores, sample, t = cpu_throughput(builder, kw, mesh, target_s=15.0, steps=max(1, min(a.steps, 3)))
        line = {"impl": "reference", "metric": "k-points/sec (H(k)+eigh)", "value": v, "unit": "k-points/s", "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "k-points/s", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": v, "unit": "k-points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import pysktb_b200 as tb
    from pysktb_b200 import _lib
    from pysktb_b200.hamiltonian import host_buffer

    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            del os.environ["NCCL_DEBUG"]               # keep stdout to the one JSON line: at VERSION and WARN level NCCL
                                                        # prints its version banner to stdout on the first communicator
        # NCCL writes its version banner to stdout when the first communicator is created (depending on the box's
        # NCCL_DEBUG): route fd 1 to stderr until the communicator exists, so that stdout carries only the JSON line
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    _, ham = W.instantiate(spec, tb, tb.scaling, device=local)
    lib = _lib.lib()

    # global mesh [mesh0*world, mesh1, mesh2]; this rank's contiguous flat slice
    gmesh = [mesh[0] * world, mesh[1], mesh[2]]
    first, count = rank * per_gpu_k, per_gpu_k
    out = torch.empty((N, count), dtype=torch.float64, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        ham.solve_kmesh(gmesh, soc=spec["soc"], first=first, count=count, out=out, check=False)

    for _ in range(max(a.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.sktb_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms = []
    barrier()
    ev0.record()
    for _ in range(a.steps):
        step()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.sktb_launch_count() - launches0
    ms = ev0.elapsed_time(ev1) / a.steps
    km, nl = _lib.C.c_double(), _lib.C.c_int32()
    _lib.check(lib.sktb_last_solve_time(ham._plan, _lib.C.byref(km), _lib.C.byref(nl)))
    kern_ms = km.value                                             # CUDA events around the last step's kernels, on their stream
    t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    value = per_gpu_k * world / (ms * 1e-3)
    checksum = float(out[:, :: max(1, count // 4096)].sum().item())
    ql_failures = ham.ql_failures()                                  # matrices whose QL did not converge (must be 0)

    # ---- e2e through the public API with host buffers -------------------------------------------------
    e2e = None
    if not a.no_e2e:
        import psutil
        avail = psutil.virtual_memory().available / max(world, 1)
        e_mesh = None
        for cand in (mesh, [mesh[0] // 2, mesh[1], mesh[2]], [mesh[0] // 4, mesh[1], mesh[2]], [max(1, mesh[0] // 8), mesh[1], mesh[2]],
                     [max(1, mesh[0] // 32), mesh[1], mesh[2]]):
            nk_e = cand[0] * cand[1] * cand[2]
            if nk_e * (24 + 8 * N) < 0.30 * avail and nk_e * (24 + 8 * N) < 48e9:
                e_mesh = cand
                break
        if e_mesh is not None:
            nk_e = e_mesh[0] * e_mesh[1] * e_mesh[2]
            kh = host_buffer((nk_e, 3), np.float64)
            with torch.cuda.device(local):
                kd = torch.empty((nk_e, 3), dtype=torch.float64, device="cuda")
                nk3 = _lib.as_i32(e_mesh)
                _lib.check(lib.sktb_kmesh(_lib.ptr_i32(nk3), 0, nk_e, kd.data_ptr(), None))
                torch.cuda.synchronize()
                kh[...] = kd.cpu().numpy()
                del kd
            vals = host_buffer((N, nk_e), np.float64)
            ham.solve_kpath(kh, soc=spec["soc"], out=vals)            # warm-up (also allocates the staging buffers)
            barrier()
            t0 = time.perf_counter()
            e_steps = max(1, min(a.steps, 2))
            for _ in range(e_steps):
                ham.solve_kpath(kh, soc=spec["soc"], out=vals)
            torch.cuda.synchronize()
            te = (time.perf_counter() - t0) / e_steps
            t_e = torch.tensor([te], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
            te = float(t_e.item())
            same = bool(np.array_equal(vals[:, :1024], out[:, :1024].cpu().numpy())) if (e_mesh[1:] == gmesh[1:] and rank == 0 and world == 1) else None
            e2e = {"value": nk_e * world / te, "unit": "k-points/s", "h2d_bytes_per_step": nk_e * 24, "d2h_bytes_per_step": nk_e * N * 8,
                   "ms_per_step": te * 1e3, "mesh_per_gpu": e_mesh, "api": "Hamiltonian.solve_kpath(ndarray, out=) -> sktb_solve_host",
                   "matches_device_path": same}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, pms = _lib.C.c_double(), _lib.C.c_double()
    _lib.check(lib.sktb_fp64_peak(local, _lib.C.byref(peak), _lib.C.byref(pms)))
    achieved = flops_per_k(N) * per_gpu_k / (kern_ms * 1e-3) / 1e12
    # DRAM traffic of the dominant kernel per launch, from the committed ncu --set full capture of the same launch shape
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tfile):
        tj = json.load(open(tfile))
        for key, rec in tj.items():
            if key in ham.last_kernel_info() and rec.get("matrix_size") == N:
                traffic = {"bytes_per_launch": rec["dram_bytes_per_launch"], "k_points_per_launch": rec["k_points_per_launch"],
                           "kernel": key, "source": rec["source"]}
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak = json.load(open(peaks_file))["hbm_gbs"] if os.path.exists(peaks_file) else 6650.0
    hbm_ach = per_gpu_k * 8.0 * N / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "fp64", "kernel": ham.last_kernel_info(), "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s",
                "frac": achieved / peak.value, "peak_source": "DFMA microbenchmark measured live on this device (MEASURED_PEAKS.json has no FP64 entry)",
                "flops_per_unit": flops_per_k(N), "kernel_ms": kern_ms, "kernel_share_of_step": kern_ms / ms, "traffic": traffic,
                "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak, "bytes_per_unit": 8 * N,
                        "peak_source": "MEASURED_PEAKS.json" if os.path.exists(peaks_file) else "fallback 6.65 TB/s"}}
    cpu = None
    if not a.no_cpu and world == 1:
        v, cores, sample, _ = cpu_throughput(builder, kw, mesh, target_s=12.0)
        cpu = {"value": v, "unit": "k-points/s", "cores": cores, "kind": "port", "sample": sample}
    line = {"metric": "k-points/sec (H(k)+eigh)", "value": value, "unit": "k-points/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "checksum": checksum, "ql_failures": ql_failures}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
