#!/bin/bash
O=gpurun_out
python -m pytest tests/test_parity_kkt.py -m gpu -x -q -k "independent_launch" 2>&1 | tail -3 > $O/r1i_pytest.log
LQ_CONFIGS="0:400,1:4" timeout 150 python tools/lq_time.py 2664 64 > $O/r1i_lq_time_2664.log 2>&1
tail -2 $O/r1i_pytest.log; cat $O/r1i_lq_time_2664.log
