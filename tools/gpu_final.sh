#!/bin/bash
# One GPU call: parity tests, the bench (both arms) and the ncu launch list of the bench command.
# (Full captures: ncu --set full --clock-control none --import-source on -k regex:k_kkt_factor_solve -s 1 -c 1 python tools/ncu_target.py 444 1;
#  for k_lq_step: CB200_LQ_LOCKSTEP=1 ... -k regex:k_lq_step -s 7 -c 1 python bench.py --steps 1 --warmup 1 --check-every 4 --no-cpu --no-single)
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/r1j_pytest.log
python bench.py > $O/r1j_bench.json 2> $O/r1j_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r1j_ref.json 2> $O/r1j_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1j_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-single > $O/r1j_ncu_bench.log 2>&1
tail -2 $O/r1j_pytest.log; cut -c1-400 $O/r1j_bench.json; cat $O/r1j_bench.err | tail -5
