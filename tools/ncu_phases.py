"""Stall samples of an ncu report grouped by the device function (line range) they fall in: where a kernel's warps spend time.
python tools/ncu_phases.py rep.ncu-rep      (needs -lineinfo and --import-source on)"""
import csv
import re
import subprocess
import sys
import os

rep = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
# function start lines of the device headers
starts = {}
for fn in ("device_core.h", "device_newton.h"):
    lst = []
    for i, line in enumerate(open(os.path.join(root, "calipso_b200", "csrc", fn)), 1):
        m = re.match(r"^(?:template <[^>]*>\s*)?(?:CB_DEVN?|__device__ __forceinline__|__device__ __noinline__)\s+[\w:<>\*& ]*?\b(\w+)\(", line)
        if m:
            lst.append((i, m.group(1)))
    starts[fn] = lst


def func_of(fn, ln):
    name = "(top)"
    for s, n in starts.get(fn, []):
        if s <= ln:
            name = n
        else:
            break
    return name


groups = {"J v (jacobian_times, sparse_dot)": {"jacobian_times", "sparse_dot", "sparse_dot4", "group_sum", "residual_error"},
          "reduced solve (ldl_solve_smem, chain runs, sweeps, TMA parts)": {"ldl_solve_smem", "ldl_solve", "chain_run_forward", "chain_run_backward",
                                                                             "sweep_forward_blocked", "sweep_backward_blocked", "part_issue", "part_wait",
                                                                             "part_prefetch", "named_bar_sync", "named_bar_arrive", "mbar_wait", "chain_task", "load_phase"},
          "factorisation (ldl_factor, supernodes, panel, GEMM tiles)": {"ldl_factor", "factor_supernode_big", "factor_supernode", "panel_factor_tc", "panel_factor",
                                                                         "dmma_8x8x4", "for_each_supernode", "ksrc_load", "kkt_entries", "factorize_regularized",
                                                                         "inertia_correction"},
          "rhs / recovery / refinement bookkeeping": {"reduced_rhs", "recover_step", "direction_symmetric", "iterative_refinement", "search_direction", "gmres_fallback"},
          "callbacks, cone, residual, merit, line search (device_newton.h + reductions)": set()}
cur, hdr, tot = None, None, {}
allsmp = 0.0
byfn, why = {}, {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if cur and hdr and len(r) > 8 and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        try:
            smp = float(d["# Samples"])
        except ValueError:
            continue
        f = func_of(cur, int(r[0]))
        g = next((k for k, v in groups.items() if f in v), "callbacks, cone, residual, merit, line search (device_newton.h + reductions)")
        tot[g] = tot.get(g, 0.0) + smp
        allsmp += smp
        byfn[f] = byfn.get(f, 0.0) + smp
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k:
                try:
                    why.setdefault(f, {})[k[6:]] = why.setdefault(f, {}).get(k[6:], 0.0) + float(v)
                except ValueError:
                    pass
print(f"# {rep}: warp stall samples by device function group ({allsmp:.0f} samples)")
for g, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{100 * v / allsmp:5.1f} %  {g}")
print("# by device function (top 24), with the three most frequent stall reasons of its samples")
for f, v in sorted(byfn.items(), key=lambda kv: -kv[1])[:24]:
    top = " ".join(f"{k}:{100 * x / max(v, 1):.0f}%" for k, x in sorted(why.get(f, {}).items(), key=lambda kv: -kv[1])[:3])
    print(f"{100 * v / allsmp:5.1f} %  {f:28s} {top}")
