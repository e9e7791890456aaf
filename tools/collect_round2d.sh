#!/bin/bash
# End-of-round evidence, part d (round 2, third session, after the FP64-issue diet of the scalar chains): everything the final state changes -
# tests, smoke, bench lines of every config, reference arm, secondary throughput, size scan, launch list of the T step, ncu --set full
# (+ source page) of the three kernels of the T step.
set -x
O=gpurun_out; mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r2d_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2d_smoke.log 2>&1
python bench.py --steps 3 --warmup 3 > $O/r2d_bench_T.json 2> $O/r2d_bench_T.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2d_bench_T_reference.json 2> /dev/null
for w in C1 C2 C3 C4 C5; do python bench.py --workload $w --no-cpu --steps 2 --warmup 3 2>/dev/null | tail -1 > $O/r2d_bench_$w.json; done
{ python tools/bench_vec.py T 65536; python tools/bench_vec.py C4 8192; python tools/bench_dos.py T 64 64 64 400; python tools/bench_dos.py T 32 32 32 400 ldos;
  python tools/bench_dos.py C4 28 28 28 400 ldos; python tools/bench_dos.py T 20 20 20 400 dos auto; python tools/bench_dos.py C4 10 10 10 400 dos auto; python tools/bench_latency.py; } > $O/r2d_secondary.log 2>&1
{ python tools/bench_heev.py 8,12,16,24,32,40,48,56,64,72,96,128,160 8192; python tools/bench_heev.py 192,256,384,512 592; python tools/bench_heev.py 32,64 8192 vec; } > $O/r2d_nscan.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2d_launches_T.csv python bench.py --steps 2 --warmup 3 --mesh 256 256 256 --no-cpu --no-e2e --no-secondary > /dev/null 2>&1
full() { name=$1; kn=$2; skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$kn -s $skip -c 1 -o $O/r2d_full_$name -f "$@" > /dev/null 2>&1
  ncu -i $O/r2d_full_$name.ncu-rep --page raw --csv > $O/r2d_full_$name.raw.csv 2>/dev/null
  ncu -i $O/r2d_full_$name.ncu-rep --page source --csv > $O/r2d_src_$name.csv 2>/dev/null
  rm -f $O/r2d_full_$name.ncu-rep; }
full pair_wide32 pair_wide32 2 python tools/run_solve.py T 64 64 64 3
full pair_tail16 pair_tail16 2 python tools/run_solve.py T 64 64 64 3
full tql_small tql_small 2 python tools/run_solve.py T 64 64 64 3
full cta64 tridiag_cta64 1 python tools/run_solve.py C4 32 32 32 2
cat $O/r2d_pytest.log; tail -3 $O/r2d_smoke.log
