#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/r1k_pytest.log
python bench.py > $O/r1k_bench.json 2> $O/r1k_bench.err
tail -2 $O/r1k_pytest.log; cut -c1-300 $O/r1k_bench.json; tail -3 $O/r1k_bench.err
