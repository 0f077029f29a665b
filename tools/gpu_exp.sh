#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/r1l_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r1l_smoke.log 2>&1
tail -2 $O/r1l_pytest.log; tail -1 $O/r1l_smoke.log
