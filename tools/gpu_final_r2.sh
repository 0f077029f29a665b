#!/bin/bash
# End-of-round-2 GPU call (one GPU): parity tests, smoke, the bench (both arms), the ncu launch list of the bench command and a
# full capture of one launch of every kernel (summarised by tools/ncu_summary.py).
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/r2f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2f_smoke.log 2>&1
python bench.py > $O/r2f_bench.json 2> $O/r2f_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2f_ref.json 2> $O/r2f_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2f_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-literal > $O/r2f_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:^k_ -o $O/r2f_all_kernels -f python tools/ncu_all_kernels.py 444 > $O/r2f_all_kernels.log 2>&1
ncu -i $O/r2f_all_kernels.ncu-rep --page raw --csv > $O/r2f_all_kernels_raw_full.csv 2>/dev/null
python tools/ncu_summary.py $O/r2f_all_kernels_raw_full.csv > $O/r2f_all_kernels_summary.txt 2>&1
rm -f $O/r2f_all_kernels.ncu-rep
tail -n 2 $O/r2f_pytest.log; tail -n 1 $O/r2f_smoke.log; cut -c1-300 $O/r2f_bench.json; tail -n 3 $O/r2f_bench.err; cat $O/r2f_all_kernels_summary.txt
python tools/r2_diff_probe.py calipso_b200/libcalipso_b200.so 2>&1 | tail -n 1
