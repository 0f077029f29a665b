#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/r1g_pytest.log
timeout 200 compute-sanitizer --tool memcheck python -m pytest tests/test_scatter.py -m gpu -x -q > $O/r1g_sanitizer_scatter.log 2>&1
python tools/scatter_time.py 1332 > $O/r1g_scatter_time.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/r1g_smoke.log 2>&1
tail -2 $O/r1g_pytest.log; tail -4 $O/r1g_sanitizer_scatter.log; cat $O/r1g_scatter_time.log; tail -1 $O/r1g_smoke.log
