#!/bin/bash
O=gpurun_out; mkdir -p $O
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 2>&1 | grep -v "^\[W\|Warning\|warn" | tail -4 | tee $O/mg_dist_check.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 > $O/mg_bench_T_2gpu.json 2> $O/mg_bench_T_2gpu.err
tail -c 1500 $O/mg_bench_T_2gpu.json; tail -3 $O/mg_bench_T_2gpu.err
