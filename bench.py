#!/usr/bin/env python
"""bench.py -- Newton iterations/s of CALIPSO's Newton/KKT hot path on B200 (BASELINE.json metric).

One "step" = one complete solve! (src/solver/solve.jl:8-377) of a batch of independent cfg3-shaped instances
(LQC(40,36,12,100), N = 4584, synthetic, seeded) resident on the GPU: every Newton iteration runs residual assembly,
KKT assembly, supernodal LDL^T with inertia correction, refinement, cone search and the filter line search on the
device.  value = Newton iterations completed by all instances of all ranks / device time (max over ranks).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

Under torchrun (N > 1) every rank owns `--batch` instances (weak scaling, no data-path collective; one NCCL
all-reduce of convergence counts per check).  `--impl reference` times the CPU oracle (the restated reference
algorithm, reference schedule) on the host cores instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from calipso_b200 import lqc  # noqa: E402

LOCKSTEP = os.environ.get("CB200_LQ_LOCKSTEP", "0") not in ("", "0")   # one k_lq_step launch per Newton iteration (A/B)

CFG = dict(T=40, n_x=36, n_u=12, n_soc=100)       # BASELINE.json configs[3] (cfg3)
METRIC = "newton_iterations_per_second"
UNIT = "Newton it/s"


def make_instances(count, first_seed):
    return [lqc.cfg3(first_seed + i) for i in range(count)]


# ------------------------------------------------------------------------------------------------- CPU oracle legs
def _oracle_worker(args):
    seeds, reference_schedule, budget_s = args
    from oracle import oracle as orc
    iters = 0
    t0 = time.perf_counter()
    solves = 0
    kkt_ms = []
    for s in seeds:
        P = lqc.cfg3(s)
        o = orc.from_problem(P, options=dict(reference_schedule=reference_schedule))
        o.use_superlu_fallback()
        o.initialize(P.x0)
        rc = o.solve()
        iters += o.stats["total_iterations"] - 1
        solves += 1
        if time.perf_counter() - t0 > budget_s:
            break
    return iters, time.perf_counter() - t0, solves


def oracle_throughput(cores, seeds, reference_schedule, budget_s):
    """Newton it/s of the oracle on `cores` host processes (one instance at a time per process)."""
    import multiprocessing as mp
    chunks = [seeds[i::cores] for i in range(cores)]
    t0 = time.perf_counter()
    if cores == 1:
        res = [_oracle_worker((chunks[0], reference_schedule, budget_s))]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_oracle_worker, [(c, reference_schedule, budget_s) for c in chunks])
    wall = time.perf_counter() - t0
    iters = sum(r[0] for r in res)
    solves = sum(r[2] for r in res)
    return iters / wall, iters, solves, wall


def oracle_kkt_solve_ms(seed=3000):
    """One numeric factorisation + one solve of the reduced system on one core (the 'KKT solve' unit)."""
    from oracle import oracle as orc
    P = lqc.cfg3(seed - 3000)
    o = orc.from_problem(P)
    o.initialize(P.x0)
    o.solve_begin()
    for _ in range(3):
        if o.newton_iteration() == 2:
            o.outer_update()
    o.evaluate(2 | 16 | 32)
    o.cone_eval(barrier=True, barrier_gradient=True)
    o.residual_eval()
    o.evaluate(64 | 128 | 256)
    o.cone_eval(jacobian=True)
    o.set_scalars(eps_p=1e-7, eps_d=1e-7)
    o.residual_jacobian_variables()
    o.residual_jacobian_variables_symmetric()
    ts = []
    for _ in range(20):
        t0 = time.perf_counter()
        o.factorize()
        o.search_direction_symmetric(factorize=False)
        ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = max(cores, 8)
    vals = []
    for step in range(args.warmup + args.steps):
        seeds = list(range(step * per_step, (step + 1) * per_step))
        v, iters, solves, wall = oracle_throughput(cores, seeds, 1, 60.0)
        if step >= args.warmup:
            vals.append((v, wall))
    value = float(np.mean([v for v, _ in vals]))
    ms = 1e3 * float(np.mean([w for _, w in vals]))
    sample = f"{per_step} complete solve! runs of cfg3 per step on {cores} processes (one instance per process at a time), reference schedule (>=3 factorisations per Newton step)"
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                data="synthetic",
                config=dict(workload="cfg3 LQC(T=40,n_x=36,n_u=12,n_soc=100) N=4584 total=8496, complete solve! per instance",
                            instances_per_step=per_step),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                note="CPU oracle = C restatement of the reference algorithm (the Julia reference cannot run here; "
                     "ordering and LU fallback are stand-ins, see oracle/oracle.h)")
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------- GPU leg
class ClockSampler:
    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from calipso_b200 import _lib
    from calipso_b200.solver import BatchKKT

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout (one JSON line only)
        saved_stdout0 = os.dup(1)                       # ... and whatever the communicator set-up still prints
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            t0 = torch.zeros(1, device="cuda")
            dist.all_reduce(t0)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout0, 1)
            os.close(saved_stdout0)
    B = args.batch
    distinct = min(B, args.distinct)
    # instances: `distinct` different seeds per rank, tiled over the batch (values differ per seed, pattern shared)
    Ps = make_instances(distinct, first_seed=rank * distinct)
    plist = [Ps[i % distinct] for i in range(B)]
    k = BatchKKT(Ps[0], batch=B, device=local)
    info = k.info()
    k.load_lq(plist)
    X0 = np.stack([P.x0 for P in plist])
    if world > 1:
        saved_stdout = os.dup(1)                        # anything NCCL prints while initialising goes to stderr
        os.dup2(2, 1)
        try:
            uid = [k.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            k.comm_init(rank, world, uid[0])
            k.allreduce_counts()                        # first collective (lazy communicator setup) inside the redirect
            t0 = torch.zeros(1, device="cuda")
            dist.all_reduce(t0)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    stream = torch.cuda.ExternalStream(k.lib.cb200_stream(k.h), device=torch.device("cuda", local))

    def barrier():
        k.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn, reps):
        """device time of `reps` calls of fn on the handle's stream (ms per call), max over ranks"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / reps
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    launches = [0, 0]              # all kernels of this library, k_lq_step launches
    lq_ms = [0.0]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def one_solve():
        k.initialize(X0)           # H2D of the primal guesses is part of a solve! call (initialize!, initialize.jl:9)
        k.lq_begin()
        ev[0].record(stream)
        r = k.lq_solve(max_steps=args.max_newton, check_every=args.check_every)   # syncs at every convergence check
        ev[1].record(stream)
        ev[1].synchronize()
        lq_ms[0] += ev[0].elapsed_time(ev[1])
        checks = (r["steps"] + args.check_every - 1) // args.check_every
        lq = r["steps"] if LOCKSTEP else checks     # fused: one k_lq_step launch carries a whole check interval
        launches[0] += 1 + lq + checks              # k_lq_begin + k_lq_step + k_count_states
        launches[1] += lq
        return r

    for _ in range(args.warmup):
        r = one_solve()
    st = k.stats()
    iters_per_solve = int((st["total_iterations"] - 1).sum())
    if world > 1:
        t = torch.tensor([iters_per_solve], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        total_iters = int(t.item())
    else:
        total_iters = iters_per_solve
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches[:] = [0, 0]
    lq_ms[0] = 0.0
    st0 = {kk: v.copy() for kk, v in k.stats().items()}
    ms_step = timed(one_solve, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    value = total_iters / (ms_step * 1e-3)
    gpu_launches = launches[0]
    conv = k.allreduce_counts()
    stats = {kk: v for kk, v in k.stats().items()}

    # ---- dominant kernel of the step: k_lq_step (one Newton iteration of every running instance per launch).
    # Algorithmic bytes = SURVEY.md section 8(d): per instance B_newton = B_res + B_cone + n_trials (B_asm + B_factor)
    # + (1 + n_refine) (B_solve + B_spmv), with the factorisation / solve counts the kernel actually performed.
    P0 = Ps[0]
    nW, nG, nC = len(P0.W_rowval), len(P0.G_rowval), len(P0.C_rowval)
    N_, T_, nK, nL = info["N"], info["total"], info["nnzK"], info["nnzL"]
    b_asm = 12 * (nW + nG + nC) + 8 * nK
    b_factor = 12 * nK + 12 * nL + 16 * N_
    b_solve = 24 * nL + 24 * N_
    b_spmv = 12 * ((2 * nW - info["n"]) + 2 * nG + 2 * nC) + 16 * T_
    b_res = 8 * (2 * T_ + 3 * info["n"] + 2 * info["m"] + info["p"])
    b_cone = 80 * info["p"]
    d_fact = int((stats["factorizations"] - st0["factorizations"]).sum())
    d_solv = int((stats["solves"] - st0["solves"]).sum())
    newton_bytes = d_fact * (b_asm + b_factor) + d_solv * (b_solve + b_spmv) + args.steps * iters_per_solve * (b_res + b_cone)
    n_lq_launches = launches[1]
    ms_lq_launch = lq_ms[0] / max(n_lq_launches, 1)
    # ---- dominant kernel: KKT factor + solve (assemble + LDL^T + one reduced solve with recovery), roofline vs HBM
    k.lq_begin()
    k.lq_step(4)                   # realistic interior point
    k.set_scalars(eps_p=1e-7, eps_d=1e-7)
    for _ in range(3):
        k.kkt_factor_solve(1)
    ms_kkt = timed(lambda: k.kkt_factor_solve(1), 10)
    ms_fact = timed(lambda: k.kkt_factor_solve(0), 10)                 # assemble + factor only
    ms_solve = (timed(lambda: k.kkt_factor_solve(5), 5) - ms_fact) / 5  # one reduced solve (+ rhs / recovery)
    b_unit = 12 * info["nnzK"] + 36 * info["nnzL"] + 40 * info["N"]
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "of fallback: 6.65 TB/s (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        try:
            measured = float(json.load(open(peaks_path)).get("hbm_gbs") or 0.0)
        except (ValueError, TypeError, OSError):
            measured = 0.0
        if measured > 0.0:
            peak, peak_src = measured, "of measured: MEASURED_PEAKS.json hbm_gbs (copy bandwidth)"
    kkt_achieved = b_unit * B / (ms_kkt * 1e-3) / 1e9
    tpath2 = os.path.join(ROOT, "profiles", "lq_step_traffic.json")
    # measured DRAM bytes of one Newton iteration of one instance (ncu capture of a launch in which every instance ran
    # exactly one iteration), scaled to the Newton iterations an average launch of the timed region carried
    traffic = None
    if os.path.exists(tpath2):
        per_it = json.load(open(tpath2)).get("dram_bytes_per_instance_iteration")
        if per_it:
            traffic = per_it * args.steps * iters_per_solve / max(n_lq_launches, 1)
    tpath = os.path.join(ROOT, "profiles", "kkt_factor_solve_traffic.json")
    kkt_traffic = json.load(open(tpath)).get("dram_bytes_per_instance") * B if os.path.exists(tpath) else None   # capture at batch 444, scaled
    achieved = newton_bytes / max(n_lq_launches, 1) / (ms_lq_launch * 1e-3) / 1e9
    roofline = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic,
                    kernel="k_lq_step", launches_in_timed_region=n_lq_launches, ms_per_launch=ms_lq_launch,
                    algorithmic_bytes_per_launch=newton_bytes / max(n_lq_launches, 1),
                    algorithmic_bytes_formula="SURVEY 8(d): n_fact*(B_asm+B_factor) + n_solves*(B_solve+B_spmv) + n_newton*(B_res+B_cone)",
                    factorizations=d_fact, reduced_solves=d_solv, peak_source=peak_src,
                    newton_iterations_per_launch=args.steps * iters_per_solve / max(n_lq_launches, 1),
                    note="launch duration = CUDA events around cb200_lq_solve on the handle's stream / k_lq_step launches "
                         "(includes the per-check 32-byte counter read-back); one launch carries up to check_every Newton "
                         "iterations of every instance, a CTA leaves when its instance has converged; traffic = ncu DRAM "
                         "bytes per instance-iteration (profiles/lq_step_traffic.json) x iterations per launch",
                    kkt_solve=dict(kernel="k_kkt_factor_solve", achieved=kkt_achieved, frac=kkt_achieved / peak,
                                   algorithmic_bytes_per_kkt_solve=b_unit, kkt_solves_per_launch=B, ms_per_launch=ms_kkt,
                                   kkt_solve_ms_per_instance_batched=ms_kkt / B, traffic=kkt_traffic,
                                   factor=dict(ms_per_launch=ms_fact, algorithmic_bytes=b_asm + b_factor,
                                               achieved=(b_asm + b_factor) * B / (ms_fact * 1e-3) / 1e9,
                                               frac=(b_asm + b_factor) * B / (ms_fact * 1e-3) / 1e9 / peak),
                                   reduced_solve=dict(ms_per_launch=ms_solve, algorithmic_bytes=b_solve,
                                                      achieved=b_solve * B / (ms_solve * 1e-3) / 1e9,
                                                      frac=b_solve * B / (ms_solve * 1e-3) / 1e9 / peak,
                                                      what="L, D, L' sweeps of the reduced system incl. reduced rhs and "
                                                           "recovery: B_solve = 24 nnz(L) + 24 N"),
                                   what="assemble + LDL' factor + 1 reduced solve with recovery per instance: the 'KKT solve' "
                                        "unit of SURVEY 8(d), B_unit = 12 nnz(K) + 36 nnz(L) + 40 N"))
    # ---- e2e: the reference-facing hot path through the C ABI with HOST buffers (pinned), copies inside the timing
    k.lq_begin()
    k.lq_step(4)
    k.lq_evaluate(2 | 16 | 32)
    k.cone(barrier=True, barrier_gradient=True, product=True)
    names_in = ["POINT", "DUAL", "SCALARS", "GRADIENT", "EQ_DUAL_GRAD", "CONE_DUAL_GRAD", "EQUALITY", "CONE", "W_VALUES",
                "G_VALUES", "C_VALUES"]
    names_out = ["STEP", "CANDIDATE", "SCALARS"]
    host_in = {}
    for nm in names_in:
        a = k.get(nm)
        t = torch.empty(a.shape, dtype=torch.float64).pin_memory()
        t.numpy()[...] = a
        host_in[nm] = t
    host_out = {nm: torch.empty((B, k.length(nm)), dtype=torch.float64).pin_memory() for nm in names_out}
    stats_out = torch.empty((B, _lib.I_COUNT), dtype=torch.int32).pin_memory()
    h2d = sum(t.numel() * 8 for t in host_in.values())
    d2h = sum(t.numel() * 8 for t in host_out.values()) + stats_out.numel() * 4
    lib, A = k.lib, _lib.A
    # the batch is split over a few handles (each with its own stream) so that the H2D copies of one chunk overlap the
    # kernels and the D2H copies of the others; every chunk does the complete upload -> hot path -> read-back
    from calipso_b200.sharding import shard_range
    nchunks = max(1, min(args.e2e_chunks, B))
    bounds = [shard_range(B, c, nchunks) for c in range(nchunks)]
    chunk_handles = [BatchKKT(Ps[0], batch=e - b, device=local) for b, e in bounds]

    def at(t, b0, ctype):
        return _lib.C.cast(t.data_ptr() + b0 * t.shape[1] * t.element_size(), ctype)

    def e2e_step():
        for (b0, e0_), kc in zip(bounds, chunk_handles):
            n_, hc = e0_ - b0, kc.h
            for nm, t in host_in.items():
                lib.cb200_set_array(hc, A[nm], at(t, b0, _lib.c_dp), 0, n_)
            lib.cb200_cone(hc, 7, 0)
            lib.cb200_residual(hc)
            lib.cb200_search_direction(hc)
            lib.cb200_cone_search(hc)
            for nm, t in host_out.items():
                lib.cb200_get_array_async(hc, A[nm], at(t, b0, _lib.c_dp), 0, n_)
            lib.cb200_get_stats_async(hc, at(stats_out, b0, _lib.c_ip), 0, n_)
        for kc in chunk_handles:
            lib.cb200_synchronize(kc.h)

    for _ in range(2):
        e2e_step()
    assert int(stats_out[:, _lib.I["status"]].abs().sum()) == 0, "e2e step reported solver errors"
    ms_e2e = timed(e2e_step, max(3, args.steps))
    e2e_value = B * world / (ms_e2e * 1e-3)
    e2e = dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=ms_e2e,
               handles_per_gpu=nchunks,
               what="one Newton step per instance through the C ABI from pinned host buffers: H2D of evaluate!'s outputs + "
                    "point, cone!, residual!, search_direction!, cone search, D2H of step/candidate/scalars/stats; the batch "
                    "is split over handles_per_gpu handles (one stream each) so copies overlap kernels")
    for kc in chunk_handles:
        kc.close()

    extras = {}
    if rank == 0 and world == 1 and not args.no_single:
        # cfg3 literal: ONE instance on one B200 (latency-bound), and cfg4 literal: 8 instances per GPU
        for bb, key in ((1, "cfg3_single_instance"), (8, "cfg4_8_instances_per_gpu")):
            kk = BatchKKT(Ps[0], batch=bb, device=local)
            kk.load_lq([Ps[i % distinct] for i in range(bb)])
            x0 = np.stack([Ps[i % distinct].x0 for i in range(bb)])
            st2 = torch.cuda.ExternalStream(kk.lib.cb200_stream(kk.h), device=torch.device("cuda", local))

            def solve2():
                kk.initialize(x0)
                kk.lq_begin()
                kk.lq_solve(max_steps=args.max_newton, check_every=args.check_every)
            for _ in range(2):
                solve2()
            kk.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st2)
            for _ in range(3):
                solve2()
            e1.record(st2)
            kk.synchronize()
            ms = e0.elapsed_time(e1) / 3
            its = int((kk.stats()["total_iterations"] - 1).sum())
            kk.lq_begin()
            kk.lq_step(4)
            kk.set_scalars(eps_p=1e-7, eps_d=1e-7)
            kk.kkt_factor_solve(1)
            kk.synchronize()
            e0.record(st2)
            for _ in range(10):
                kk.kkt_factor_solve(1)
            e1.record(st2)
            kk.synchronize()
            extras[key] = dict(batch=bb, newton_it_per_s=its / (ms * 1e-3), ms_per_solve=ms, newton_iterations=its,
                               kkt_factor_solve_ms=e0.elapsed_time(e1) / 10)
            kk.close()

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            v, iters, solves, wall = oracle_throughput(1, list(range(64)), 1, args.cpu_budget)
            v2, iters2, solves2, wall2 = oracle_throughput(1, list(range(64)), 0, args.cpu_budget / 2)
            cpu = dict(value=v, unit=UNIT, cores=1, kind="port",
                       sample=f"{solves} complete solve! runs of cfg3 (seeds 0..{solves - 1}), {iters} Newton iterations, "
                              f"{wall:.1f} s on 1 core, reference schedule (>=3 factorisations per step)",
                       deduplicated_schedule_value=v2, kkt_factor_solve_ms=oracle_kkt_solve_ms())
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                    data="synthetic",
                    config=dict(workload=f"cfg4-style batch of cfg3 LQC(T=40,n_x=36,n_u=12,n_soc=100) instances, "
                                         f"{B} per GPU, one complete solve! of every instance per step",
                                N=info["N"], total=info["total"], n=info["n"], m=info["m"], p=info["p"],
                                nnz_K_upper=info["nnzK"], nnz_L=info["nnzL"], supernodes=info["supernodes"],
                                levels=info["levels"], batch_per_gpu=B, distinct_seeds_per_gpu=distinct,
                                newton_iterations_per_step=total_iters,
                                newton_iterations_per_launch_max=args.check_every,
                                schedule="lock-step (one launch per Newton iteration)" if LOCKSTEP else
                                         "independent (a launch carries up to check_every iterations of an instance)",
                                l2_policy="inputs larger than L2: per-GPU working set = batch x ~4.5 MB",
                                parallelism=f"instances sharded {B}/GPU, NCCL all-reduce of convergence counts only"),
                    e2e=e2e, gpu_launches=gpu_launches, roofline=roofline, cpu_baseline=cpu, clocks=clocks,
                    converged=conv, fallbacks=int(stats["fallbacks"].sum()), **extras)
        print(json.dumps(line))
    k.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1332, help="instances per GPU (9 per SM: three waves of 3 resident CTAs x 148 SMs)")
    ap.add_argument("--distinct", type=int, default=64, help="distinct seeds per GPU, tiled over the batch")
    ap.add_argument("--max-newton", type=int, default=400)
    ap.add_argument("--check-every", type=int, default=400,
                    help="Newton iterations per k_lq_step launch / convergence check (default: run every instance to "
                         "convergence inside one launch)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-single", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--e2e-chunks", type=int, default=8, help="handles (streams) the e2e step pipelines the batch over")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
