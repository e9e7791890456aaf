#!/bin/bash
O=gpurun_out; mkdir -p $O
# launch list of the C5 edge LDOS (ribbon N=512, 592 k, atoms 0 1 510 511)
cat > /tmp/c5ldos.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch, pysktb_b200 as tb
from pysktb_b200 import workloads as W
_, ham = W.instantiate(W.zigzag_ribbon(256), tb, tb.scaling)
sg = tb.SurfaceGreensFunction(ham, [0, 1, 510, 511])
E = np.linspace(-3, 3, 200)
k = np.stack([np.linspace(0, 1, 592, endpoint=False), np.zeros(592), np.zeros(592)], axis=1)
for rep in range(2): d = sg.edge_dos(E, k_points=k, eta=0.05, soc=False)
torch.cuda.synchronize(); print(ham.last_kernel_info()[:100])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/ts13_launches_c5ldos.csv python /tmp/c5ldos.py > $O/ts13_ncu.log 2>&1
python tools/launch_breakdown.py $O/ts13_launches_c5ldos.csv | tee $O/ts13_breakdown.log
