#!/bin/bash
# One GPU call: parity tests, the bench (both arms), the ncu launch list of the bench command and the full captures.
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/r1f_pytest.log
python bench.py > $O/r1f_bench.json 2> $O/r1f_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r1f_ref.json 2> $O/r1f_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1f_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-single > $O/r1f_ncu_bench.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_kkt_factor_solve -s 1 -c 1 -f -o $O/r1f_kkt \
    python tools/ncu_target.py 444 1 > $O/r1f_ncu_kkt.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_lq_step -s 7 -c 1 -f -o $O/r1f_lqstep \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-single > $O/r1f_ncu_lqstep.log 2>&1
tail -2 $O/r1f_pytest.log; cut -c1-300 $O/r1f_bench.json
