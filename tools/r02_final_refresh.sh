#!/bin/bash
# refresh of the artefacts the last kernel change (stein) touches: GPU tests, smoke, C5 bench line, C5 secondary numbers, C5 LDOS launch list
O=gpurun_out; mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r2b_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2b_smoke.log 2>&1
python bench.py --workload C5 --no-cpu --steps 3 --warmup 3 2>/dev/null | tail -1 > $O/r2b_bench_C5.json
{ python tools/bench_vec.py C5 592; python tools/bench_dos.py C5 2048 1 1 200 ldos; python tools/bench_dos.py C5 4096 1 1 400; } > $O/r2b_secondary.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2b_launches_C5ldos.csv python tools/bench_dos.py C5 2048 1 1 200 ldos > /dev/null 2>&1
cat $O/r2b_pytest.log; tail -2 $O/r2b_smoke.log; python -c "
import json; d=json.load(open('$O/r2b_bench_C5.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'] is not None, d['secondary']['edge_ldos']['value'])"
