#!/bin/bash
# session 3: launch lists of the dense-inverse DOS (C4: CTA kernel, T: warp kernel) + ncu --set full of both kernels
O=gpurun_out; mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/s3_launches_gf_C4.csv python tools/bench_dos.py C4 10 10 10 400 dos inverse > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/s3_launches_gf_T.csv python tools/bench_dos.py T 20 20 20 400 dos inverse > /dev/null 2>&1
full() { name=$1; kn=$2; skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$kn -s $skip -c 1 -o $O/s3_full_$name -f "$@" > /dev/null 2>&1
  ncu -i $O/s3_full_$name.ncu-rep --page raw --csv > $O/s3_full_$name.raw.csv 2>/dev/null
  ncu -i $O/s3_full_$name.ncu-rep --page source --csv > $O/s3_src_$name.csv 2>/dev/null
  rm -f $O/s3_full_$name.ncu-rep; }
full gf_cta gf_inverse_cta 1 python tools/bench_dos.py C4 10 10 10 400 dos inverse
full gf_warp gf_inverse_warp 1 python tools/bench_dos.py T 20 20 20 400 dos inverse
python tools/launch_breakdown.py $O/s3_launches_gf_C4.csv; python tools/launch_breakdown.py $O/s3_launches_gf_T.csv
ls -la $O/s3_*
