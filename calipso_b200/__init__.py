"""calipso_b200 -- CALIPSO.jl's per-iteration Newton/KKT hot path on NVIDIA B200 (sm_100a).

Public surface (mirrors src/solver of the reference): Solver, solve (solve!), initialize (initialize!), Options,
LDLSolver / ldl_solver (LinearSolver seam), BatchKKT (batched handle), lqc (synthetic LQ-conic instances).
"""
from .solver import BatchKKT, LDLSolver, Options, Solver, initialize, ldl_solver, solve  # noqa: F401
