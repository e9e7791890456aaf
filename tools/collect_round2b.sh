#!/bin/bash
# End-of-round evidence, part b (round 2, after the two-stage reduction): tests, smoke, T and C5 bench lines, size scan of the large-N paths,
# launch lists and ncu --set full captures of the new kernels.  Everything lands in gpurun_out/r2b_*.
set -x
O=gpurun_out
mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r2b_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2b_smoke.log 2>&1
python bench.py --steps 3 --warmup 3 > $O/r2b_bench_T.json 2> $O/r2b_bench_T.err
python bench.py --workload C5 --no-cpu --steps 3 --warmup 3 2>/dev/null | tail -1 > $O/r2b_bench_C5.json
{ python tools/bench_heev.py 160,192,208,224,256,320,384,448,512 592; python tools/bench_heev.py 512 4096; SKTB_TWOSTAGE_MIN=100000 python tools/bench_heev.py 208,224,256,320,384,512 592;
  python tools/bench_heev.py 128,256 296 vec; python tools/bench_heev.py 512 148 vec; TS_TIME=592 python tools/test_twostage.py 256,384,512 2 | grep "stage1.*ms"; } > $O/r2b_nscan.log 2>&1
{ python tools/bench_vec.py C5 592; python tools/bench_dos.py C5 2048 1 1 200 ldos; python tools/bench_dos.py C5 4096 1 1 400; } > $O/r2b_secondary.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2b_launches_C5.csv python tools/run_solve.py C5 888 1 1 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2b_launches_C5ldos.csv python tools/bench_dos.py C5 2048 1 1 200 ldos > /dev/null 2>&1
full() { # name kernel-regex skip command...
  name=$1; kn=$2; skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$kn -s $skip -c 1 -o $O/r2b_full_$name -f "$@" > /dev/null 2>&1
  ncu -i $O/r2b_full_$name.ncu-rep --page raw --csv > $O/r2b_full_$name.raw.csv 2>/dev/null
  if [ "$name" = "stage1" ] || [ "$name" = "stage2" ]; then ncu -i $O/r2b_full_$name.ncu-rep --page source --csv > $O/r2b_src_$name.csv 2>/dev/null; fi
  rm -f $O/r2b_full_$name.ncu-rep
}
full stage1 stage1_kernel 1 python tools/bench_heev.py 512 296
full stage2 stage2_kernel 1 python tools/bench_heev.py 512 296
full bisect_prod bisect_prod 1 python tools/bench_heev.py 512 296
full stein_ts stein_kernel 1 python tools/bench_dos.py C5 592 1 1 200 ldos
full rows_ts rows_kernel 1 python tools/bench_dos.py C5 592 1 1 200 ldos
ls -la $O/r2b_* | tail -30
