"""Timings of the same library under two shared-memory plans (CB200_PLAN), in separate processes:
python tools/r2_plan_ab.py PLAN[:SOLVE_BATCH] ..."""
import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
for spec in sys.argv[1:] or ["1", "0"]:
    plan, _, sb = spec.partition(":")
    env = dict(os.environ, CB200_PLAN=plan)
    print(f"--- CB200_PLAN={plan} solve batch {sb or 1332}", flush=True)
    subprocess.run([sys.executable, os.path.join(here, "r2_ab.py"), "--a", "none", "--solve-batch", sb or "1332", "--distinct", "64"], env=env)
