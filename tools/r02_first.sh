#!/bin/bash
# First GPU call of round 2: tests, smoke, short bench lines, launch list and source-level ncu of the headline kernels.
set -x
O=gpurun_out
mkdir -p $O
nproc; free -g | head -2; nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/r02a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02a_smoke.log 2>&1
python bench.py --steps 2 --warmup 3 > $O/r02a_bench_T.json 2> $O/r02a_bench_T.err
python bench.py --workload C4 --no-e2e --no-cpu --steps 1 --warmup 1 > $O/r02a_bench_C4.json 2> $O/r02a_bench_C4.err
python bench.py --workload C5 --no-e2e --no-cpu --steps 1 --warmup 1 > $O/r02a_bench_C5.json 2> $O/r02a_bench_C5.err
for kn in pair_wide32 pair_tail16 tql_small; do
  ncu --set full --clock-control none --import-source on -k regex:$kn -s 2 -c 1 -o $O/r02a_full_$kn -f python tools/run_solve.py T 64 64 64 3 > /dev/null 2>&1
  ncu -i $O/r02a_full_$kn.ncu-rep --page source --csv > $O/r02a_src_$kn.csv 2>/dev/null
  ncu -i $O/r02a_full_$kn.ncu-rep --page raw --csv > $O/r02a_raw_$kn.csv 2>/dev/null
  rm -f $O/r02a_full_$kn.ncu-rep
done
ls -la $O/r02a_*
