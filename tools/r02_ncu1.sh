#!/bin/bash
O=gpurun_out; mkdir -p $O
for kn in pair_wide32 pair_tail16; do
  ncu --set full --clock-control none --import-source on -k regex:$kn -s 2 -c 1 -o $O/r02b_full_$kn -f python tools/run_solve.py T 64 64 64 3 > /dev/null 2>&1
  ncu -i $O/r02b_full_$kn.ncu-rep --page source --csv > $O/r02b_src_$kn.csv 2>/dev/null
  ncu -i $O/r02b_full_$kn.ncu-rep --page raw --csv > $O/r02b_raw_$kn.csv 2>/dev/null
  rm -f $O/r02b_full_$kn.ncu-rep
done
ncu --set full --clock-control none --import-source on -k regex:vecbt_big -s 1 -c 1 -o $O/r02b_full_vecbt -f python tools/bench_heev.py 512 148 vec > /dev/null 2>&1
ncu -i $O/r02b_full_vecbt.ncu-rep --page source --csv > $O/r02b_src_vecbt.csv 2>/dev/null
ncu -i $O/r02b_full_vecbt.ncu-rep --page raw --csv > $O/r02b_raw_vecbt.csv 2>/dev/null
rm -f $O/r02b_full_vecbt.ncu-rep
ls -la $O/r02b_*
