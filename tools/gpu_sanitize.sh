#!/bin/bash
# compute-sanitizer passes over the GPU parity tests (round 2): memcheck on the new entry points and on the cfg3-size
# search direction, racecheck on the barrier-free solve chain (named barriers + TMA) at small sizes.
O=gpurun_out/r2_sanitizer.txt
: > $O
run() { echo "# $*" >> $O; timeout 900 "$@" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|error" | head -20 >> $O; }
run compute-sanitizer --tool memcheck python -m pytest tests/test_stage_scatter.py tests/test_scatter.py tests/test_golden.py tests/test_differentiate.py tests/test_warmstart.py -m gpu -q -x
run compute-sanitizer --tool memcheck python -m pytest tests/test_full_size.py -m gpu -q -x -k "search_direction or linear_solver"
run compute-sanitizer --tool racecheck python -m pytest tests/test_golden.py -m gpu -q -x -k "step_tiny or solve_tiny_0"
cat $O
