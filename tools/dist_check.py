"""torchrun check: multi-GPU results == single-GPU results (DOS all-reduce, eigenvalue all-gather).
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dist_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import pysktb_b200 as tb
from pysktb_b200 import workloads as W

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
spec = W.sb_buckled()
_, ham = W.instantiate(spec, tb, tb.scaling, device=local)
gf = tb.GreensFunction(ham)
E = np.linspace(-2.5, 2.5, 101)
multi = gf.dos(E, nk=[37, 29, 1], eta=0.05)                      # sharded + all-reduced
ld_multi = gf.ldos(E, atom_indices=[1, 0], nk=[11, 13, 1], eta=0.05)
k = np.random.default_rng(0).random((1001, 3))
ev_multi = tb.dist.solve_kpath_sharded(ham, list(k))
dist.barrier()
# single-rank reference: temporarily hide the process group from the library
import pysktb_b200.greens as G
orig = G._dist
G._dist = lambda: None
single = gf.dos(E, nk=[37, 29, 1], eta=0.05)
ld_single = gf.ldos(E, atom_indices=[1, 0], nk=[11, 13, 1], eta=0.05)
G._dist = orig
ev_single = ham.solve_kpath(list(k))
ok = (np.allclose(multi, single, rtol=1e-12, atol=0) and np.allclose(ld_multi, ld_single, rtol=1e-12, atol=0)
      and np.array_equal(ev_multi, ev_single))
print(f"rank {rank}/{world}: dos rel diff {np.abs(multi/single-1).max():.2e}, ldos {np.abs(ld_multi/ld_single-1).max():.2e}, "
      f"evals identical {np.array_equal(ev_multi, ev_single)} -> {'OK' if ok else 'FAIL'}")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
