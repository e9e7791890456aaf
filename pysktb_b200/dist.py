"""k-point sharding across GPUs (one process per GPU, torch.distributed; NCCL over NVLink on a B200 box).

The path shards with no data-path exchange: every k-point is independent.  Each rank solves
a contiguous slice of the k list / flat mesh index; the only collectives are the all-gather of eigenvalues
(when the caller wants the whole array on every rank) and the all-reduce of DOS histograms
(pysktb_b200.greens).  On CPU-only hosts the same code runs over the gloo backend with the solver
replaced by a caller-supplied function (tests/test_dist_gloo.py) - the product solver itself is CUDA-only.
"""
import numpy as np

from .greens import shard_range  # noqa: F401  (re-export)


def world():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def solve_kpath_sharded(ham, k_list, soc=True, gather=True, solver=None):
    """Each rank solves its contiguous slice of `k_list`; with gather=True every rank returns the full
    (N, nk) eigenvalue array (all_gather of padded slices), else its local (N, nk_local) slice and (lo, hi).

    `solver(k_slice) -> (N, nk_local) ndarray` defaults to `ham.solve_kpath(k_slice, soc=soc)`."""
    import torch
    import torch.distributed as dist

    rank, size = world()
    k = np.asarray([np.asarray(x, dtype=float) for x in k_list], dtype=float).reshape(-1, 3)
    lo, hi = shard_range(len(k), rank, size)
    solver = solver or (lambda ks: ham.solve_kpath(list(ks), soc=soc))
    local = np.asarray(solver(k[lo:hi])) if hi > lo else None
    if size == 1:
        return local if gather else (local, (lo, hi))
    if not gather:
        return local, (lo, hi)
    N = local.shape[0] if local is not None else None
    # agree on N (a rank with an empty slice does not know it)
    use_cuda = dist.get_backend() == "nccl"
    # the plan's device, not torch's current one: the plan picks its GPU from LOCAL_RANK (hamiltonian.py) and the caller
    # need not have called torch.cuda.set_device
    dev = torch.device("cuda", _plan_device(ham)) if use_cuda else torch.device("cpu")
    n_t = torch.tensor([N or 0], dtype=torch.int64, device=dev)
    dist.all_reduce(n_t, op=dist.ReduceOp.MAX)
    N = int(n_t.item())
    width = -(-len(k) // size)
    pad = torch.zeros((N, width), dtype=torch.float64, device=dev)
    if local is not None:
        pad[:, : hi - lo] = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    parts = [torch.empty_like(pad) for _ in range(size)]
    dist.all_gather(parts, pad)
    out = np.empty((N, len(k)))
    for r, part in enumerate(parts):
        a, b = shard_range(len(k), r, size)
        out[:, a:b] = part[:, : b - a].cpu().numpy()
    return out


def _plan_device(ham):
    import torch

    d = getattr(ham, "_device", None)
    return int(d) if d is not None else torch.cuda.current_device()


def allreduce_sum(arr, device=None):
    """Sum a host array over ranks (DOS partial histograms).  `device`: CUDA device index (or a Hamiltonian) the NCCL
    collective runs on; defaults to torch's current device."""
    import torch
    import torch.distributed as dist

    rank, size = world()
    if size == 1:
        return arr
    use_cuda = dist.get_backend() == "nccl"
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if use_cuda:
        if device is not None and not isinstance(device, int):
            device = _plan_device(device)
        t = t.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
