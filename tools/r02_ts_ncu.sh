#!/bin/bash
# ncu --set full of the two-stage kernels (N = 512, 296 matrices); raw + source CSVs come back, the reports stay on the box
O=gpurun_out; mkdir -p $O
TAG=${1:-ts}
for kn in stage1_kernel stage2_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kn -s 1 -c 1 -o $O/${TAG}_full_$kn -f python tools/bench_heev.py 512 296 > $O/${TAG}_ncu_$kn.log 2>&1
  ncu -i $O/${TAG}_full_$kn.ncu-rep --page source --csv > $O/${TAG}_src_$kn.csv 2>/dev/null
  ncu -i $O/${TAG}_full_$kn.ncu-rep --page raw --csv > $O/${TAG}_raw_$kn.csv 2>/dev/null
  rm -f $O/${TAG}_full_$kn.ncu-rep
done
ls -la $O/${TAG}_*
