#!/bin/bash
# End-of-round evidence: tests, smoke, bench lines for every config, secondary throughput numbers, ncu launch list
# and full captures of every kernel family.  Everything lands in gpurun_out/final_*.
set -x
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1
python bench.py --impl reference --steps 1 --warmup 0 > $O/final_bench_T_reference.json 2> /dev/null
python bench.py --steps 3 --warmup 3 > $O/final_bench_T.json 2> $O/final_bench_T.err
for w in C1 C2 C3 C4 C5; do python bench.py --workload $w --no-e2e --no-cpu --steps 2 --warmup 3 2>/dev/null | tail -1 > $O/final_bench_$w.json; done
{ python tools/bench_vec.py T 65536; python tools/bench_vec.py C3 65536; python tools/bench_vec.py C4 8192; python tools/bench_vec.py C5 148;
  python tools/bench_dos.py C2 4096 4096 1 400; python tools/bench_dos.py T 64 64 64 400; python tools/bench_dos.py T 32 32 32 400 ldos;
  python tools/bench_dos.py C4 24 24 24 400; python tools/bench_dos.py C4 28 28 28 400 ldos; python tools/bench_dos.py C5 592 1 1 200 ldos;
  python tools/bench_latency.py; } > $O/final_secondary.log 2>&1
{ python tools/bench_heev.py 8,16,24,26,32,33,40,48,56,64,65,72,80,96,111,128,160 8192; python tools/bench_heev.py 192,256,384,512;
  python tools/bench_heev.py 32,64 8192 vec; python tools/bench_heev.py 96,128,256 296 vec; } > $O/final_nscan.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/final_launches_T.csv python bench.py --steps 2 --warmup 3 --mesh 256 256 256 --no-cpu --no-e2e > $O/final_ncu_bench.log 2>&1
for kn in pair_wide32 pair_tail16 tql_small; do ncu --set full --clock-control none --import-source on -k regex:$kn -s 2 -c 1 -o $O/final_full_$kn -f python tools/run_solve.py T 64 64 64 3 > /dev/null 2>&1; done
ncu --set full --clock-control none --import-source on -k regex:tridiag_cta64 -s 1 -c 1 -o $O/final_full_cta64 -f python tools/run_solve.py C4 32 32 32 2 > /dev/null 2>&1
for kn in tqlrec vecapply32 vecbt32; do ncu --set full --clock-control none --import-source on -k regex:$kn -s 1 -c 1 -o $O/final_full_$kn -f python tools/bench_vec.py T 65536 > /dev/null 2>&1; done
for kn in tqlrec vecapply64 vecbt64; do ncu --set full --clock-control none --import-source on -k regex:$kn -s 1 -c 1 -o $O/final_full_${kn}_n64 -f python tools/bench_vec.py C4 8192 > /dev/null 2>&1; done
ncu --set full --clock-control none --import-source on -k regex:heev_generic -s 1 -c 1 -o $O/final_full_generic512 -f python tools/run_solve.py C5 296 1 1 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lorentz_partial -s 1 -c 1 -o $O/final_full_lorentz -f python tools/bench_dos.py C2 2048 2048 1 400 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:solve_n2 -s 1 -c 1 -o $O/final_full_n2 -f python tools/run_solve.py C2 4096 4096 1 2 > /dev/null 2>&1
for r in $O/final_full_*.ncu-rep; do ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null; rm -f $r; done   # reports are ~10 MB each: keep the raw metric tables only (gpurun_out is capped at 64 MiB)
ls -la $O/final_* | tail -40
