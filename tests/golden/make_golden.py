"""Generate the committed golden fixtures of the Newton/KKT path from the CPU oracle (oracle/: the C restatement of the
reference; the Julia reference itself cannot run in this environment, see oracle/oracle.h).

  python tests/golden/make_golden.py        # rewrites tests/golden/*.npz

Each `step_*.npz` holds one Newton step of a seeded LQ-conic instance: the state entering residual!/search_direction!
(point, duals, scalars, callback outputs) and what the oracle computes from it (residual, direction, inertia, trial /
refinement counts, cone-search halvings, candidate).  `solve_*.npz` holds complete solve! runs (final point, iteration
counts).  The elimination order is stored with each fixture because everything downstream of it is compared exactly.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from calipso_b200 import lqc  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from test_parity_kkt import oracle_at_iteration  # noqa: E402
import backends  # noqa: E402
from calipso_b200.solver import BatchKKT  # noqa: E402

STEP_CASES = [("tiny", 0, 0), ("tiny", 0, 4), ("tiny", 1, 7), ("tiny", 2, 2)]
SOLVE_CASES = [("tiny", 0), ("tiny", 3), ("cfg2", 0), ("rocket", 0)]


def instance(name, seed):
    """LQ-conic instance by fixture name: the seeded LQC family, or the reference's rocket-landing example."""
    if name == "rocket":
        import problems
        return problems.rocket_landing(seed)
    return getattr(lqc, name)(seed)


def product_perm(P):
    """The product's elimination order (host symbolic analysis; the emulation library runs the same code)."""
    k = BatchKKT(P, binding=backends.binding("emul"))
    perm, _, _ = k.symbolic()
    return perm


def main():
    for name, seed, iters in STEP_CASES:
        P = getattr(lqc, name)(seed)
        perm = product_perm(P)
        o = oracle_at_iteration(P, iters, perm=perm)
        o.evaluate(2 | 16 | 32)
        o.cone_eval(barrier=True, barrier_gradient=True)
        o.merit_gradient_eval()
        o.residual_eval()
        o.evaluate(64 | 128 | 256)
        o.cone_eval(jacobian=True)
        sc = o.scalars()
        state = dict(perm=perm, point=o.solution.copy(), dual=o.dual.copy(), kappa=sc["kappa"], tau=sc["tau"], rho=sc["rho"],
                     eps_p_last=sc["eps_p_last"], gradient=o.gradient.copy(), eq_dual_grad=o.eq_dual_grad.copy(),
                     cone_dual_grad=o.cone_dual_grad.copy(), equality=o.equality.copy(), cone=o.cone.copy(),
                     W_val=o.W_val.copy(), G_val=o.G_val.copy(), C_val=o.C_val.copy(), residual=o.residual.copy())
        rc = o.search_direction()
        assert rc == 0
        st = o.stats
        out = dict(step=o.step.copy(), inertia=np.array(o.inertia), n_trials=st["n_trials"], n_refine=st["n_refine"],
                   refine_ok=st["refine_ok"], used_lu=st["used_lu"], eps_p=o.scalars()["eps_p"], eps_d=o.scalars()["eps_d"])
        assert o.cone_search() == 0
        out.update(k_s=o.stats["k_s"], k_t=o.stats["k_t"], candidate=o.candidate.copy())
        np.savez_compressed(os.path.join(HERE, f"step_{name}_{seed}_{iters}.npz"), **state, **out)
    for name, seed in SOLVE_CASES:
        P = instance(name, seed)
        perm = product_perm(P)
        o = orc.from_problem(P, perm=perm)
        o.use_superlu_fallback()
        o.initialize(P.x0)
        assert o.solve() == 1
        np.savez_compressed(os.path.join(HERE, f"solve_{name}_{seed}.npz"), perm=perm, solution=o.solution.copy(),
                            total_iterations=o.stats["total_iterations"], outer=o.stats["outer"],
                            lu_fallbacks=o.stats["lu_fallbacks"])
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
