#!/bin/bash
# session 3: 2-GPU bench line (multi-GPU equality check now includes the dense-inverse DOS) + the gloo-free NCCL check script
O=gpurun_out; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 > $O/s3_bench_T_2gpu.json 2> $O/s3_bench_T_2gpu.err
tail -c 1500 $O/s3_bench_T_2gpu.err; python -c "
import json; d=json.loads(open('$O/s3_bench_T_2gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d.get('multi_gpu_check'), {k:v.get('value') for k,v in d['secondary'].items()})"
