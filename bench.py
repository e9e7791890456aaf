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
            CUDA-event time of the kernels, against a DFMA microbenchmark measured live on the same device
            (MEASURED_PEAKS.json has no FP64 entry; the method is on record in profiles/peaks_fp64.json); the HBM
            side (8N bytes per k out) is reported beside it.
  secondary the consumers of the same solver that BASELINE.json's configs name: C4 -> LDOS with eigenvectors
            consumed on the device at 256^3, C5 -> edge-projected LDOS at 4096 k, T -> DOS on a fixed global
            256^3 x 400 E mesh with the NCCL all-reduce of the histogram INSIDE the timed region (strong scaling).
  multi_gpu_check (N > 1, outside the timed region) sharded DOS / LDOS / all-gathered eigenvalues against the same
            calls on one rank: {dos_rel, ldos_rel, evals_equal}.
  cpu_baseline / --impl reference   the CPU arm, always in a FRESH subprocess (no CUDA context, OMP/OpenBLAS threads
            pinned to 1 in the environment before numpy is imported, one process per host core): the numpy oracle
            port of the reference's path (kind "port", the line's value: the strong, conservative baseline) and -
            when build() staged the reference's own modules into oracle/_ref/ - the reference as shipped,
            Hamiltonian.solve_kpath(parallel=1) (joblib over all cores), beside it ("reference_as_shipped").
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (workloads builder, kwargs, per-GPU mesh)   - the sizes BASELINE.json / SURVEY.md §8(d) state
    "T": ("ce_spdf_1atom", {}, [512, 512, 512]),
    "C1": ("cubic_sp3", {}, [256, 256, 256]),
    "C2": ("graphene_pz", {}, [4096, 4096, 1]),
    "C3": ("sb_buckled", {}, [1024, 1024, 1]),
    "C4": ("ce_spdf_2atom", {}, [256, 256, 256]),
    "C5": ("zigzag_ribbon", {"n_chains": 256}, [4096, 1, 1]),
}


def flops_per_k(N, vectors=False):
    return (16.0 if vectors else 16.0 / 3.0) * N ** 3


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
# CPU arm.  Runs ONLY inside the worker subprocess (`bench.py --cpu-worker ...`): a fresh interpreter whose
# environment pins the BLAS/OpenMP pools to one thread before numpy is imported, with no CUDA context, no torch and
# no pinned buffers to fork from.  One process per host core; every timed pass is the SAME fixed sample.
# ------------------------------------------------------------------------------------------------------
_ORC = None


def _cpu_init(spec_name, kw):
    global _ORC
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


def _worker_port(spec_name, kw, mesh, steps, warmup, step_s):
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    total = mesh[0] * mesh[1] * mesh[2]
    with mp.get_context("fork").Pool(cores, initializer=_cpu_init, initargs=(spec_name, kw)) as pool:
        def one_pass(chunk):
            t0 = time.perf_counter()
            pool.map(_cpu_chunk, [(mesh, (c * chunk) % total, (c * chunk) % total + chunk) for c in range(cores)])
            return time.perf_counter() - t0
        one_pass(2)                                            # imports, model build in every worker
        cal_chunk = 8
        t = one_pass(cal_chunk)
        while t < 1.0 and cal_chunk * cores * 4 <= total:      # calibrate on >= 1 s of work, not on 16 points
            cal_chunk *= 4
            t = one_pass(cal_chunk)
        chunk = int(max(cal_chunk, min(total // cores, round(cal_chunk * step_s / t))))
        if chunk >= 1024:
            chunk -= chunk % 256
        for _ in range(warmup):
            one_pass(chunk)
        times = [one_pass(chunk) for _ in range(steps)]
    n = chunk * cores
    tm = sum(times) / len(times)
    return {"value": n / tm, "unit": "k-points/s", "cores": cores, "kind": "port", "steps": steps, "warmup": warmup, "ms_per_step": tm * 1e3,
            "timed_s": sum(times), "step_times_s": [round(x, 3) for x in times],
            "sample": f"{n} points of the {mesh[0]}x{mesh[1]}x{mesh[2]} mesh per step ({cores} processes x {chunk} consecutive k each, "
                      f"OMP/OpenBLAS threads = 1 per process), the numpy oracle port of hamiltonian.py:98-126,275-317"}


def _worker_reference(spec_name, kw, mesh, steps, warmup, step_s):
    """The reference as shipped: its own Hamiltonian(..., numba=0).solve_kpath(k_list, parallel=1), i.e. joblib/loky
    over cpu_count() processes (hamiltonian.py:166-189), imported from the modules build() staged into oracle/_ref/."""
    os.environ["SKTB_REFERENCE_ROOT"] = os.path.join(ROOT, "oracle", "_ref")
    import contextlib
    import io
    import numpy as np
    from oracle import ref_loader
    from pysktb_b200 import workloads as W
    if not ref_loader.available():
        return {"unavailable": "oracle/_ref/pysktb not staged (build() stages it where /root/reference exists)"}
    sys.path.insert(0, os.environ["SKTB_REFERENCE_ROOT"])          # loky workers re-import pysktb.* (a namespace package there)
    os.environ["PYTHONPATH"] = os.environ["SKTB_REFERENCE_ROOT"] + os.pathsep + os.environ.get("PYTHONPATH", "")
    ref = ref_loader.load()
    import types
    from pysktb.hamiltonian import Hamiltonian as RefHamiltonian    # the module-level class (picklable by reference)
    ns = types.SimpleNamespace(Lattice=ref.Lattice, Atom=ref.Atom, Structure=ref.Structure, Hamiltonian=RefHamiltonian)
    spec = W.ALL[spec_name](**kw)
    with contextlib.redirect_stdout(io.StringIO()):
        _, ham = W.instantiate(spec, ns, ref.scaling, numba=0)
    cores = os.cpu_count() or 1
    par = 1 if spec["soc"] else 0            # parallel=1 with soc=False crashes in the reference: serial loop then
    from oracle.sktb_oracle import OracleModel

    def run(n):
        ks = [OracleModel.kmesh_point(mesh, i) for i in range(n)]
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            out = ham.solve_kpath(ks, eig_vectors=False, soc=spec["soc"], parallel=par)
        assert np.isfinite(out).all()
        return time.perf_counter() - t0
    run(cores * 2)                                             # loky spawn + imports
    n_cal = cores * 8
    t = run(n_cal)
    n = int(max(n_cal, min(mesh[0] * mesh[1] * mesh[2], round(n_cal * step_s / t))))
    for _ in range(warmup):
        run(n)
    times = [run(n) for _ in range(steps)]
    tm = sum(times) / len(times)
    return {"value": n / tm, "unit": "k-points/s", "cores": cores if par else 1, "kind": "reference", "steps": steps, "warmup": warmup,
            "ms_per_step": tm * 1e3, "timed_s": sum(times),
            "sample": f"first {n} points of the {mesh[0]}x{mesh[1]}x{mesh[2]} mesh per step through the reference's own "
                      f"Hamiltonian(numba=0).solve_kpath(parallel={par})" + (f" (joblib/loky, {cores} processes)" if par else " (serial: parallel=1 crashes for soc=False)")}


def cpu_worker_main(argv):
    a = json.loads(argv[0])
    fn = _worker_reference if a["mode"] == "reference" else _worker_port
    print("CPU_WORKER_RESULT " + json.dumps(fn(a["builder"], a["kw"], a["mesh"], a["steps"], a["warmup"], a["step_s"])))


def cpu_leg(mode, builder, kw, mesh, steps, warmup, step_s, timeout=900):
    """Run the CPU arm in a fresh interpreter (see above) and parse its result."""
    env = dict(os.environ)
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        env[v] = "1"
    env["CUDA_VISIBLE_DEVICES"] = ""
    for v in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(v, None)
    spec = json.dumps({"mode": mode, "builder": builder, "kw": kw, "mesh": mesh, "steps": steps, "warmup": warmup, "step_s": step_s})
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-worker", spec], env=env, capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return {"unavailable": f"cpu worker ({mode}) exceeded {timeout} s"}
    for line in r.stdout.splitlines():
        if line.startswith("CPU_WORKER_RESULT "):
            return json.loads(line[len("CPU_WORKER_RESULT "):])
    return {"unavailable": f"cpu worker ({mode}) failed rc={r.returncode}: {r.stderr.strip()[-300:]}"}


# ------------------------------------------------------------------------------------------------------
def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--cpu-worker":
        return cpu_worker_main(sys.argv[2:])
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="T", choices=list(WORKLOADS))
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--mesh", type=int, nargs=3, default=None, help="override the per-GPU mesh")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    a = ap.parse_args()

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    builder, kw, mesh = WORKLOADS[a.workload]
    mesh = list(a.mesh) if a.mesh else list(mesh)
    from pysktb_b200 import workloads as W
    spec = W.ALL[builder](**kw)
    N = len([o for _, _, orbs in spec["atoms"] for o in orbs]) * (2 if spec["soc"] else 1)
    per_gpu_k = mesh[0] * mesh[1] * mesh[2]
    config = {"workload": f"{a.workload}: {spec['name']}, N={N}, soc={spec['soc']}, eigenvalues only (value / e2e), "
                          f"uniform mesh {mesh[0]}x{mesh[1]}x{mesh[2]} per GPU ({per_gpu_k} k-points/GPU); the Hermiticity gate of "
                          f"hamiltonian.py:112 is decided once per model from the hopping tables (sktb_plan_asym_bound) and evaluated per k only "
                          f"when it can fire",
              "k_points_per_gpu": per_gpu_k, "matrix_size": N, "parallelism": f"k-sharded x{a.gpus}",
              "cache": "no reusable inputs: k generated in-kernel, output (8N B/k) >> L2"}

    if a.impl == "reference":
        if rank != 0:
            return
        # every step is the same bounded sample; (K + W) steps are sized to ~150 s in total
        step_s = max(2.0, min(12.0, 150.0 / max(1, a.steps + a.warmup)))
        res = cpu_leg("port", builder, kw, mesh, a.steps, a.warmup, step_s)
        if "unavailable" in res:
            print(json.dumps({"impl": "reference", "unavailable": res["unavailable"]}))
            return
        shipped = cpu_leg("reference", builder, kw, mesh, 1, 0, 10.0)
        v = res["value"]
        line = {"impl": "reference", "metric": "k-points/sec (H(k)+eigh)", "value": v, "unit": "k-points/s", "n_gpus": a.gpus,
                "steps": res["steps"], "warmup": res["warmup"], "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "k-points/s", "cores": res["cores"], "kind": res["kind"], "sample": res["sample"],
                                 "timed_s": res["timed_s"], "step_times_s": res["step_times_s"]},
                "reference_as_shipped": shipped,
                "e2e": {"value": v, "unit": "k-points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # the CPU baseline runs first, before this process owns a CUDA context or pinned memory
    cpu = None
    if not a.no_cpu and world == 1:
        res = cpu_leg("port", builder, kw, mesh, 1, 1, 12.0)
        if "unavailable" in res:
            cpu = res
        else:
            cpu = {"value": res["value"], "unit": "k-points/s", "cores": res["cores"], "kind": "port", "sample": res["sample"], "timed_s": res["timed_s"]}
            shipped = cpu_leg("reference", builder, kw, mesh, 1, 0, 10.0)
            cpu["reference_as_shipped"] = shipped

    import numpy as np
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

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def device_timed(fn, steps, warmup):
        """W untimed + K timed calls bracketed by barrier + synchronize, CUDA events on the current stream, max over ranks (ms/step)."""
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1) / steps)

    def step():
        # check=False: the per-k gate flags are not allocated; T's asym_bound is below 0.5e-9, so the gate cannot fire for it
        ham.solve_kmesh(gmesh, soc=spec["soc"], first=first, count=count, out=out, check=False)

    warmup = max(a.warmup, 3)
    for _ in range(warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.sktb_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(a.steps):
        step()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.sktb_launch_count() - launches0
    ms = max_over_ranks(ev0.elapsed_time(ev1) / a.steps)
    km, nl = _lib.C.c_double(), _lib.C.c_int32()
    _lib.check(lib.sktb_last_solve_time(ham._plan, _lib.C.byref(km), _lib.C.byref(nl)))
    kern_ms = km.value                                             # CUDA events around the last step's kernels, on their stream
    value = per_gpu_k * world / (ms * 1e-3)
    checksum = float(out[:, :: max(1, count // 4096)].sum().item())
    kernel_info = ham.last_kernel_info()
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
            te = max_over_ranks((time.perf_counter() - t0) / e_steps)
            same = bool(np.array_equal(vals[:, :1024], out[:, :1024].cpu().numpy())) if (e_mesh[1:] == gmesh[1:] and rank == 0 and world == 1) else None
            e2e = {"value": nk_e * world / te, "unit": "k-points/s", "h2d_bytes_per_step": nk_e * 24, "d2h_bytes_per_step": nk_e * N * 8,
                   "ms_per_step": te * 1e3, "mesh_per_gpu": e_mesh, "api": "Hamiltonian.solve_kpath(ndarray, out=) -> sktb_solve_host",
                   "matches_device_path": same}
            del kh, vals

    # ---- secondary legs: the consumers BASELINE.json's configs name ---------------------------------------
    peak, pms = _lib.C.c_double(), _lib.C.c_double()
    _lib.check(lib.sktb_fp64_peak(local, _lib.C.byref(peak), _lib.C.byref(pms)))
    secondary = {}
    if not a.no_secondary:
        # the throughput legs run the eigenpair route explicitly (for the f-orbital models T / C4 that is the DOS of LAPACK's
        # lower-triangle spectrum - what solve_kpath returns); the default method="auto" would take the dense inverse for them
        gf = tb.GreensFunction(ham, method="eigh")
        s_steps = max(1, min(a.steps, 2))
        if a.workload == "T":
            # DOS of T on a FIXED global 256^3 mesh x 400 energies (strong scaling): each rank reduces its k-slice on the
            # device, ncclAllReduce(sum, f64) of the [1][400] histogram inside the timed region (GreensFunction._reduce)
            dmesh, E = [256, 256, 256], np.linspace(-6.0, 6.0, 400)
            res = {}
            t = device_timed(lambda: res.__setitem__("dos", gf.dos(E, nk=dmesh, eta=0.05, soc=spec["soc"])), s_steps, 1)
            nkd = dmesh[0] * dmesh[1] * dmesh[2]
            secondary["dos"] = {"metric": "k-points/sec (H(k)+eigh+Lorentzian DOS, 400 energies)", "value": nkd / (t * 1e-3), "unit": "k-points/s",
                                "ms_per_step": t, "global_mesh": dmesh, "n_energies": 400, "n_gpus": world, "scaling": "strong",
                                "collective": "ncclAllReduce(sum, f64) of the [1][400] histogram, inside the timed region" if world > 1 else "none (1 rank)",
                                "normalisation": float(np.trapezoid(res["dos"], E)) if hasattr(np, "trapezoid") else float(np.trapz(res["dos"], E)),
                                "api": "GreensFunction(ham, method='eigh').dos(E, nk=[256,256,256], eta=0.05)"}
            # the reference's own numbers for this model (H(k) not Hermitian: dense inverse per (k, E), greens.py:70) on the
            # reference's default mesh: GreensFunction(ham).dos(E) = nk [20,20,20] x 400 energies = 3.2e6 complex inverses of order 32
            E2 = np.linspace(-6.0, 6.0, 400)
            ginv = tb.GreensFunction(ham)
            if ginv._use_inverse():
                t = device_timed(lambda: ginv.dos(E2, eta=0.05, soc=spec["soc"]), 1, 1 if a.warmup else 0)
                n_inv = (8000 // world + (1 if 8000 % world else 0)) * 400
                ach = 8.0 * N ** 3 * n_inv / (t * 1e-3) / 1e12
                secondary["dos_inverse"] = {"metric": "k-points/sec (H(k) + dense inverse per (k,E), 400 energies: the reference's DOS for non-Hermitian H)",
                                            "value": 8000 / (t * 1e-3), "unit": "k-points/s", "ms_per_step": t, "global_mesh": [20, 20, 20], "n_energies": 400,
                                            "inverses_per_s": 8000 * 400 / (t * 1e-3), "n_gpus": world, "kernel": ham.last_kernel_info(),
                                            "roofline": {"bound": "fp64", "achieved": ach, "peak": peak.value, "unit": "TFLOP/s", "frac": ach / peak.value,
                                                         "flops_per_unit": 8.0 * N ** 3, "unit_of_work": "one (k,E) inverse"},
                                            "api": "GreensFunction(ham).dos(E)  (method='auto' -> 'inverse', reference defaults nk=[20,20,20])"}
        if a.workload == "C4":
            # BASELINE config 4 as stated: 256^3 mesh WITH eigenvectors, consumed on the device (LDOS of both atoms, 400 E);
            # the k range is split over ranks like the main leg
            E = np.linspace(-6.0, 6.0, 400)
            lmesh = gmesh
            t = device_timed(lambda: gf.ldos(E, atom_indices=[0, 1], nk=lmesh, eta=0.05, soc=spec["soc"]), 1, 1 if a.warmup else 0)
            nkl = lmesh[0] * lmesh[1] * lmesh[2]
            ach = flops_per_k(N, True) * (nkl / world) / (t * 1e-3) / 1e12
            secondary["ldos"] = {"metric": "k-points/sec (H(k)+eigh with eigenvectors, consumed on device: LDOS 2 atoms x 400 E)", "value": nkl / (t * 1e-3),
                                 "unit": "k-points/s", "ms_per_step": t, "global_mesh": lmesh, "n_gpus": world, "kernel": ham.last_kernel_info(),
                                 "roofline": {"bound": "fp64", "achieved": ach, "peak": peak.value, "unit": "TFLOP/s", "frac": ach / peak.value,
                                              "flops_per_unit": flops_per_k(N, True)},
                                 "api": "GreensFunction(ham, method='eigh').ldos(E, atom_indices=[0,1], nk=[256,256,256], eta=0.05)"}
            # the same LDOS exactly as the reference computes it for this model (H(k) not Hermitian: dense inverse per (k, E),
            # greens.py:169-259 -> :70) on the reference's default mesh nk = [20,20,20] x 400 energies (3.2e6 inverses of order 64)
            ginv = tb.GreensFunction(ham)
            if ginv._use_inverse():
                t = device_timed(lambda: ginv.ldos(E, atom_indices=[0, 1], eta=0.05, soc=spec["soc"]), 1, 1 if a.warmup else 0)
                n_inv = (8000 // world + (1 if 8000 % world else 0)) * 400
                ach = 8.0 * N ** 3 * n_inv / (t * 1e-3) / 1e12
                secondary["ldos_inverse"] = {"metric": "k-points/sec (H(k) + dense inverse per (k,E), LDOS 2 atoms x 400 E: the reference's LDOS for non-Hermitian H)",
                                             "value": 8000 / (t * 1e-3), "unit": "k-points/s", "ms_per_step": t, "global_mesh": [20, 20, 20], "n_energies": 400,
                                             "inverses_per_s": 8000 * 400 / (t * 1e-3), "n_gpus": world, "kernel": ham.last_kernel_info(),
                                             "roofline": {"bound": "fp64", "achieved": ach, "peak": peak.value, "unit": "TFLOP/s", "frac": ach / peak.value,
                                                          "flops_per_unit": 8.0 * N ** 3, "unit_of_work": "one (k,E) inverse"},
                                             "api": "GreensFunction(ham).ldos(E, atom_indices=[0,1])  (method='auto' -> 'inverse', reference defaults nk=[20,20,20])"}
        if a.workload == "C5":
            # BASELINE config 5 as stated: 1-D sweep of 4096 k with the edge-projected LDOS (atoms 0, 1, 510, 511; 200 E, eta 0.08)
            sg = tb.SurfaceGreensFunction(ham, surface_atoms=[0, 1, 510, 511])
            E = np.linspace(-3.0, 3.0, 200)
            nke = per_gpu_k * world
            kpts = np.stack([np.linspace(0, 1, nke), np.zeros(nke), np.zeros(nke)], axis=1)
            t = device_timed(lambda: sg.edge_dos(E, k_points=kpts, eta=0.08, soc=spec["soc"]), 1, 1 if a.warmup else 0)
            secondary["edge_ldos"] = {"metric": "k-points/sec (H(k)+eigh+edge-projected LDOS, 4 atoms x 200 E)", "value": nke / (t * 1e-3), "unit": "k-points/s",
                                      "ms_per_step": t, "k_points": nke, "n_gpus": world, "kernel": ham.last_kernel_info(),
                                      "api": "SurfaceGreensFunction(ham, [0,1,510,511]).edge_dos(E, k_points=linspace(0,1,4096))"}

    # ---- multi-GPU results == single-GPU results (outside the timed region) ---------------------------------
    mg = None
    if world > 1:
        import pysktb_b200.greens as G
        from pysktb_b200 import dist as tbdist
        cspec = W.sb_buckled()
        _, cham = W.instantiate(cspec, tb, tb.scaling, device=local)
        cgf = tb.GreensFunction(cham)
        E = np.linspace(-2.5, 2.5, 101)
        kk = np.random.default_rng(0).random((1001, 3))
        multi = cgf.dos(E, nk=[37, 29, 1], eta=0.05)                       # sharded + all-reduced (NCCL)
        ld_multi = cgf.ldos(E, atom_indices=[1, 0], nk=[11, 13, 1], eta=0.05)
        ev_multi = tbdist.solve_kpath_sharded(cham, list(kk))               # sharded + all-gathered (NCCL)
        evT_multi = tbdist.solve_kpath_sharded(ham, list(kk), soc=spec["soc"])
        dosT_multi = tb.GreensFunction(ham, method="eigh").dos(E, nk=[24, 24, 24] if N <= 64 else [64, 1, 1], eta=0.05, soc=spec["soc"])
        inv_gf = tb.GreensFunction(ham, method="inverse") if N <= 64 else None     # dense inverse per (k, E), k-sharded + all-reduced
        inv_multi = inv_gf.dos(E[::5], nk=[6, 5, 4], eta=0.05, soc=spec["soc"]) if inv_gf else None
        barrier()
        orig = G._dist
        G._dist = lambda: None                                             # the same calls on this rank alone
        try:
            single = cgf.dos(E, nk=[37, 29, 1], eta=0.05)
            ld_single = cgf.ldos(E, atom_indices=[1, 0], nk=[11, 13, 1], eta=0.05)
            dosT_single = tb.GreensFunction(ham, method="eigh").dos(E, nk=[24, 24, 24] if N <= 64 else [64, 1, 1], eta=0.05, soc=spec["soc"])
            inv_single = inv_gf.dos(E[::5], nk=[6, 5, 4], eta=0.05, soc=spec["soc"]) if inv_gf else None
        finally:
            G._dist = orig
        ev_single = cham.solve_kpath(list(kk))
        evT_single = ham.solve_kpath(list(kk), soc=spec["soc"])
        rel = lambda x, y: float(np.abs(x / y - 1).max())
        loc = [rel(multi, single), rel(ld_multi, ld_single), rel(dosT_multi, dosT_single),
               0.0 if (np.array_equal(ev_multi, ev_single) and np.array_equal(evT_multi, evT_single)) else 1.0,
               rel(inv_multi, inv_single) if inv_gf else 0.0]
        t = torch.tensor(loc, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)                            # worst over ranks
        w = t.cpu().numpy()
        mg = {"dos_rel": float(w[0]), "ldos_rel": float(w[1]), "dos_rel_bench_model": float(w[2]), "evals_equal": bool(w[3] == 0.0),
              "dos_inverse_rel_bench_model": float(w[4]),
              "what": "C3 model: GreensFunction.dos [37,29,1] / ldos [11,13,1] sharded + ncclAllReduce vs one rank; dist.solve_kpath_sharded "
                      "(ncclAllGather) vs solve_kpath on 1001 random k, C3 and the bench model (bit-equality); bench model: DOS by the eigenpair route and "
                      "(N <= 64) by the dense inverse per (k,E) on a [6,5,4] mesh; worst over ranks"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    achieved = flops_per_k(N) * per_gpu_k / (kern_ms * 1e-3) / 1e12
    # DRAM traffic of the dominant kernel per launch, from the committed ncu --set full capture of the same launch shape
    traffic = None
    import glob
    for tfile in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")), reverse=True):
        tj = json.load(open(tfile))
        for key, rec in tj.items():
            if traffic is None and key in kernel_info and rec.get("matrix_size") == N:
                traffic = {"bytes_per_launch": rec["dram_bytes_per_launch"], "k_points_per_launch": rec["k_points_per_launch"],
                           "kernel": key, "source": rec["source"]}
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak = json.load(open(peaks_file))["hbm_gbs"] if os.path.exists(peaks_file) else 6650.0
    hbm_ach = per_gpu_k * 8.0 * N / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "fp64", "kernel": kernel_info, "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s",
                "frac": achieved / peak.value, "peak_source": "DFMA microbenchmark measured live on this device (sktb_fp64_peak; MEASURED_PEAKS.json has no "
                                                              "FP64 entry; method and round-2 value in profiles/peaks_fp64.json)",
                "flops_per_unit": flops_per_k(N), "kernel_ms": kern_ms, "kernel_share_of_step": kern_ms / ms, "traffic": traffic,
                "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak, "bytes_per_unit": 8 * N,
                        "peak_source": "MEASURED_PEAKS.json" if os.path.exists(peaks_file) else "fallback 6.65 TB/s"}}
    line = {"metric": "k-points/sec (H(k)+eigh)", "value": value, "unit": "k-points/s", "n_gpus": world, "steps": a.steps,
            "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "secondary": secondary or None, "multi_gpu_check": mg,
            "checksum": checksum, "ql_failures": ql_failures}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
