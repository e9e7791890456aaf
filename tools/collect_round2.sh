#!/bin/bash
# End-of-round evidence (round 2): tests, smoke, bench lines for every config at its stated size, secondary throughput numbers,
# K1 bandwidth, FP64 peak record, ncu launch lists and full captures of every kernel family.  Everything lands in gpurun_out/r2_*.
set -x
O=gpurun_out
mkdir -p $O
nproc > $O/r2_host.txt; free -g | head -2 >> $O/r2_host.txt; nvidia-smi -L >> $O/r2_host.txt
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r2_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1
python tools/write_peaks.py $O/r2_peaks_fp64.json > /dev/null 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_T_reference.json 2> /dev/null
python bench.py --steps 3 --warmup 3 > $O/r2_bench_T.json 2> $O/r2_bench_T.err
for w in C1 C2 C3 C4 C5; do python bench.py --workload $w --no-cpu --steps 2 --warmup 3 2>/dev/null | tail -1 > $O/r2_bench_$w.json; done
{ python tools/bench_vec.py T 65536; python tools/bench_vec.py C3 65536; python tools/bench_vec.py C4 8192; python tools/bench_vec.py C5 592;
  python tools/bench_dos.py C2 4096 4096 1 400; python tools/bench_dos.py T 64 64 64 400; python tools/bench_dos.py T 32 32 32 400 ldos;
  python tools/bench_dos.py C4 24 24 24 400; python tools/bench_dos.py C4 28 28 28 400 ldos; python tools/bench_dos.py C5 2048 1 1 200 ldos;
  python tools/bench_hk.py C4 C5 T; python tools/bench_latency.py; } > $O/r2_secondary.log 2>&1
{ python tools/bench_heev.py 4,8,12,16,24,26,32,33,40,48,56,64,65,72,80,96,111,128,160 8192; python tools/bench_heev.py 192,256,384,512;
  python tools/bench_heev.py 32,64 8192 vec; python tools/bench_heev.py 96,128,256 296 vec; python tools/bench_heev.py 512 148 vec; } > $O/r2_nscan.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_T.csv python bench.py --steps 2 --warmup 3 --mesh 256 256 256 --no-cpu --no-e2e --no-secondary > $O/r2_ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2_launches_C4.csv python tools/run_solve.py C4 64 64 64 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2_launches_C5ldos.csv python tools/bench_dos.py C5 2048 1 1 200 ldos > /dev/null 2>&1
full() { # name kernel-regex skip command...
  name=$1; kn=$2; skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$kn -s $skip -c 1 -o $O/r2_full_$name -f "$@" > /dev/null 2>&1
  ncu -i $O/r2_full_$name.ncu-rep --page raw --csv > $O/r2_full_$name.raw.csv 2>/dev/null
  if [ "$name" = "pair_wide32" ] || [ "$name" = "cta64" ]; then ncu -i $O/r2_full_$name.ncu-rep --page source --csv > $O/r2_src_$name.csv 2>/dev/null; fi
  rm -f $O/r2_full_$name.ncu-rep
}
full pair_wide32 pair_wide32 2 python tools/run_solve.py T 64 64 64 3
full pair_tail16 pair_tail16 2 python tools/run_solve.py T 64 64 64 3
full tql_small tql_small 2 python tools/run_solve.py T 64 64 64 3
full cta64 tridiag_cta64 1 python tools/run_solve.py C4 32 32 32 2
full solve_thread solve_thread 1 python tools/run_solve.py C1 128 128 128 2
full solve_thread12 solve_thread 1 python tools/run_solve.py C3 1024 1024 1 2
full hk hk_kernel 1 python tools/bench_hk.py C4
full generic512 heev_generic 1 python tools/run_solve.py C5 296 1 1 2
full stein stein_kernel 1 python tools/bench_heev.py 512 148 vec
full vecbt_big vecbt_big 1 python tools/bench_heev.py 512 148 vec
full vecbt_big256 vecbt_big 1 python tools/bench_heev.py 256 296 vec
full rowsq_big rowsq_big 1 python tools/bench_dos.py C5 592 1 1 200 ldos
for kn in tqlrec vecapply32 vecbt32; do full $kn $kn 1 python tools/bench_vec.py T 65536; done
for kn in tqlrec vecapply64 vecbt64; do full ${kn}_n64 $kn 1 python tools/bench_vec.py C4 8192; done
full lorentz lorentz_partial 1 python tools/bench_dos.py C2 2048 2048 1 400
full n2 solve_n2 1 python tools/run_solve.py C2 4096 4096 1 2
ls -la $O/r2_* | tail -60
