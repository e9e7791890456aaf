#!/bin/bash
# End-of-round evidence, part c (round 2, third session: dense-inverse Green's function): tests, smoke, T bench line with its CPU arms and
# the dos_inverse leg, reference arm, C4 / C5 bench lines, inverse-path timings, launch lists and ncu --set full captures of the K5 kernels.
set -x
O=gpurun_out; mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r2c_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c_smoke.log 2>&1
python bench.py --steps 3 --warmup 3 > $O/r2c_bench_T.json 2> $O/r2c_bench_T.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2c_bench_T_reference.json 2> /dev/null
for w in C4 C5; do python bench.py --workload $w --no-cpu --steps 2 --warmup 3 2>/dev/null | tail -1 > $O/r2c_bench_$w.json; done
{ python tools/bench_dos.py T 20 20 20 400 dos auto; python tools/bench_dos.py T 20 20 20 400 ldos auto; python tools/bench_dos.py C4 10 10 10 400 dos auto;
  python tools/bench_dos.py C4 10 10 10 400 ldos auto; python tools/bench_dos.py T 64 64 64 400 dos eigh; SKTB_GF_TILE=0 python tools/bench_dos.py C4 10 10 10 400 dos auto;
  SKTB_GF_TILE=0 SKTB_GF_WARP=0 python tools/bench_dos.py T 20 20 20 400 dos auto; SKTB_GF_TILE1=1 python tools/bench_dos.py T 20 20 20 400 dos auto; } > $O/r2c_gf_throughput.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2c_launches_gf_T.csv python tools/bench_dos.py T 20 20 20 400 dos auto > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2c_launches_gf_C4.csv python tools/bench_dos.py C4 10 10 10 400 dos auto > /dev/null 2>&1
full() { name=$1; kn=$2; skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$kn -s $skip -c 1 -o $O/r2c_full_$name -f "$@" > /dev/null 2>&1
  ncu -i $O/r2c_full_$name.ncu-rep --page raw --csv > $O/r2c_full_$name.raw.csv 2>/dev/null
  ncu -i $O/r2c_full_$name.ncu-rep --page source --csv > $O/r2c_src_$name.csv 2>/dev/null
  rm -f $O/r2c_full_$name.ncu-rep; }
full gf_warp gf_inverse_warp 1 python tools/bench_dos.py T 20 20 20 400 dos auto
full gf_tile4 gf_inverse_tile 1 python tools/bench_dos.py C4 10 10 10 400 dos auto
SKTB_GF_TILE=0 full gf_cta gf_inverse_cta 1 python tools/bench_dos.py C4 10 10 10 400 dos auto
SKTB_GF_TILE=0 SKTB_GF_WARP=0 full gf_smem "gf_inverse_kernel" 1 python tools/bench_dos.py T 20 20 20 400 dos auto
cat $O/r2c_pytest.log; tail -3 $O/r2c_smoke.log; cat $O/r2c_gf_throughput.log
python -c "
import json
for w in ('T','C4','C5'):
    d=json.load(open('$O/r2c_bench_%s.json'%w)); print(w, d['value'], d['e2e']['value'], d['roofline']['frac'], {k:(v.get('value')) for k,v in d.get('secondary',{}).items()}, d.get('cpu_baseline',{}).get('value'))
print(open('$O/r2c_bench_T_reference.json').read()[:600])"
