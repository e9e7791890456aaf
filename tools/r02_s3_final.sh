#!/bin/bash
# final check of the committed state: all GPU tests, smoke
O=gpurun_out; mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r2e_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2e_smoke.log 2>&1
cat $O/r2e_pytest.log; tail -6 $O/r2e_smoke.log
