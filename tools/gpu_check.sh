#!/bin/bash
# Short GPU check: the parity suite, the smoke test, and the batched-solve timing (tools/lq_time.py).
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/r1m_pytest.log
LQ_CONFIGS="0:400" timeout 60 python tools/lq_time.py 1332 64 > $O/r1m_lq_time.log 2>&1
tail -2 $O/r1m_pytest.log; cat $O/r1m_lq_time.log
