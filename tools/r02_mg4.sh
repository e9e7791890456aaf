#!/bin/bash
# 4-GPU hardware check: sharded DOS / LDOS / eigenvalue gather against one rank (tools/dist_check.py), then the bench line with multi_gpu_check and the DOS secondary leg
O=gpurun_out; mkdir -p $O
nvidia-smi -L | tee $O/mg4_gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 2>&1 | grep -v "^\[W\|Warning\|warn" | tail -4 | tee $O/mg4_dist_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 2 --warmup 3 --no-cpu > $O/mg4_bench_T_4gpu.json 2> $O/mg4_bench_T_4gpu.err
tail -c 1200 $O/mg4_bench_T_4gpu.json; tail -3 $O/mg4_bench_T_4gpu.err
