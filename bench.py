#!/usr/bin/env python
"""bench.py -- Newton iterations/s of CALIPSO's Newton/KKT hot path on B200 (BASELINE.json metric).

One "step" = one complete solve! (src/solver/solve.jl:8-377) of a batch of independent cfg3-shaped instances
(LQC(40,36,12,100), N = 4584, synthetic, seeded; BASELINE.json configs[3]/[4]) resident on the GPU: every Newton iteration
runs residual assembly, KKT assembly, supernodal LDL^T with inertia correction, refinement, cone search and the filter
line search on the device.  value = Newton iterations completed by all instances of all ranks / device time (max over
ranks).  e2e = the same complete solves through the C ABI from pinned HOST buffers: problem data and guesses are uploaded
and solutions read back inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

Under torchrun (N > 1) every rank owns `--batch` instances (weak scaling, no data-path collective; one NCCL all-reduce of
convergence counts per check).  `--impl reference` times the CPU oracle (the restated reference algorithm, reference
schedule) on all host cores on a bounded sample of the same workload.
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

METRIC = "newton_iterations_per_second"
UNIT = "Newton it/s"
SHAPE = dict(T=40, n_x=36, n_u=12, n_soc=100, n=1908, m=1440, p=1236, N=4584, total=8496)     # cfg3


def workload_config(batch, distinct, check_every):
    """The `config` object of both arms (the reference arm runs a bounded sample of exactly this workload)."""
    return dict(workload=f"cfg4-style batch of cfg3 LQC(T=40,n_x=36,n_u=12,n_soc=100) instances (BASELINE.json configs[3]/[4] "
                         f"shape), {batch} per GPU, one complete solve! of every instance per step",
                N=SHAPE["N"], total=SHAPE["total"], n=SHAPE["n"], m=SHAPE["m"], p=SHAPE["p"], batch_per_gpu=batch,
                distinct_seeds_per_gpu=distinct, seeds="1000*3 + rank*distinct + i (calipso_b200/lqc.py)",
                newton_iterations_per_launch_max=check_every,
                schedule="lock-step (one launch per Newton iteration)" if LOCKSTEP else
                         "independent (a launch carries up to check_every iterations of an instance)",
                l2_policy="inputs larger than L2: per-GPU working set = batch x ~4.5 MB",
                parallelism=f"instances sharded {batch}/GPU, NCCL all-reduce of convergence counts only")


# ------------------------------------------------------------------------------------------------- CPU oracle legs
_W = {}      # per-process state of the oracle workers
_NEXT = None  # shared ticket counter (created before the workers fork)


def _oracle_prepare(args):
    """(untimed) build the step's instances and their oracle solvers in this worker: generation, ordering, symbolic analysis"""
    seeds, reference_schedule = args
    from oracle import oracle as orc
    objs = []
    for s in seeds:
        P = lqc.cfg3(s)
        o = orc.from_problem(P, options=dict(reference_schedule=reference_schedule))
        o.use_superlu_fallback()
        objs.append((o, P.x0.copy()))
    _W[reference_schedule] = objs
    return len(objs)


def _oracle_solve_tickets(reference_schedule):
    """(timed) this worker pulls instance tickets from the shared counter until the step's instances are used up:
    initialize! + solve! of each; returns (Newton iterations, instances solved, busy seconds)"""
    objs = _W[reference_schedule]
    iters = solved = 0
    t0 = time.perf_counter()
    while True:
        with _NEXT.get_lock():
            i = _NEXT.value
            _NEXT.value = i + 1
        if i >= len(objs):
            break
        o, x0 = objs[i]
        o.initialize(x0)
        o.solve()
        iters += o.stats["total_iterations"] - 1
        solved += 1
    return iters, solved, time.perf_counter() - t0


class OraclePool:
    """`cores` persistent worker processes sharing a step of `instances` prepared cfg3 instances (every worker holds all of
    them; the instances of a step are handed out one at a time, so no worker idles while another still has several to do)."""

    def __init__(self, cores, instances, first_seed=0):
        import multiprocessing as mp
        global _NEXT
        self.cores, self.instances = cores, instances
        self.seeds = [first_seed + i for i in range(instances)]
        _NEXT = mp.get_context("fork").Value("i", 0)
        self.pools = [mp.get_context("fork").Pool(1) for _ in range(cores)] if cores > 1 else None
        self.ready = set()

    def prepare(self, reference_schedule):
        if reference_schedule in self.ready:
            return
        if self.pools is None:
            _oracle_prepare((self.seeds, reference_schedule))
        else:
            for r in [p.apply_async(_oracle_prepare, ((self.seeds, reference_schedule),)) for p in self.pools]:
                r.get()
        self.ready.add(reference_schedule)

    def step(self, reference_schedule):
        """one timed step; returns (Newton iterations, wall seconds, busy seconds summed over the workers)"""
        self.prepare(reference_schedule)
        _NEXT.value = 0
        t0 = time.perf_counter()
        if self.pools is None:
            res = [_oracle_solve_tickets(reference_schedule)]
        else:
            res = [r.get() for r in [p.apply_async(_oracle_solve_tickets, (reference_schedule,)) for p in self.pools]]
        wall = time.perf_counter() - t0
        assert sum(r[1] for r in res) == self.instances
        return sum(r[0] for r in res), wall, sum(r[2] for r in res)

    def close(self):
        if self.pools:
            for p in self.pools:
                p.terminate()


def oracle_kkt_solve_ms(seed=0):
    """One numeric factorisation + one solve of the reduced system on one core (the 'KKT solve' unit)."""
    from oracle import oracle as orc
    P = lqc.cfg3(seed)
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
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    instances = 4 * cores
    pool = OraclePool(cores, instances)
    pool.prepare(1)
    vals = []
    for step in range(args.warmup + args.steps):
        iters, wall, busy = pool.step(1)
        if step >= args.warmup:
            vals.append((iters / wall, wall, iters, busy))
    dedup_iters, dedup_wall, _ = pool.step(0)
    pool.close()
    value = float(np.mean([v[0] for v in vals]))
    ms = 1e3 * float(np.mean([v[1] for v in vals]))
    it_step = float(np.mean([v[2] for v in vals]))
    busy = float(np.mean([v[3] for v in vals]))
    sample = (f"{instances} complete solve! runs of cfg3 per step (seeds 0..{instances - 1}) handed out one at a time to {cores} "
              f"persistent processes; instance generation, ordering and symbolic analysis outside the timed region; reference "
              f"schedule (>=3 factorisations per Newton step, as src/solver does)")
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                data="synthetic", config=workload_config(args.batch, min(args.batch, args.distinct or args.batch), args.check_every),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample,
                                  newton_iterations_per_step=it_step,
                                  per_core_value_while_all_cores_run=it_step / busy,
                                  scheduling_efficiency=busy / (cores * ms * 1e-3),
                                  deduplicated_schedule_value=dedup_iters / dedup_wall),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                note="CPU oracle = C restatement of the reference algorithm (the Julia reference cannot run here; the AMD "
                     "ordering and the LU fallback are restated / stand-ins, see oracle/oracle.h); a step is a bounded sample "
                     "of the GPU arm's workload (same instance family, fewer instances)")
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
    from calipso_b200.sharding import shard_range
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
            import datetime
            # (short collective timeout: a rank that dies must not keep the others -- and the GPU box -- waiting for minutes)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
            t0 = torch.zeros(1, device="cuda")
            dist.all_reduce(t0)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout0, 1)
            os.close(saved_stdout0)
    B = args.batch
    distinct = min(B, args.distinct or B)
    # instances: `distinct` different seeds per rank (by default all of the batch), values differ per seed, pattern shared
    Ps = [lqc.cfg3(rank * distinct + i) for i in range(distinct)]
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
    dev = torch.device("cuda", local)
    stream = torch.cuda.ExternalStream(k.lib.cb200_stream(k.h), device=dev)

    def barrier():
        k.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def rank_max(ms):
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def rank_all(v):
        if world == 1:
            return [float(v)]
        t = torch.zeros(world, device="cuda", dtype=torch.float64)
        t[rank] = v
        dist.all_reduce(t)
        return [float(x) for x in t.tolist()]

    def rank_sum(v):
        if world > 1:
            t = torch.tensor([v], device="cuda", dtype=torch.int64)
            dist.all_reduce(t)
            v = int(t.item())
        return int(v)

    def timed(fn, reps, st=stream):
        """device time of `reps` calls of fn on stream st (ms per call), max over ranks"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            fn()
        e1.record(st)
        barrier()
        return rank_max(e0.elapsed_time(e1) / reps)

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
    total_iters = rank_sum(iters_per_solve)
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
    fallbacks_per_solve = int((stats["fallbacks"] - st0["fallbacks"]).sum()) / max(args.steps, 1)
    unrefined = int((stats["unrefined_steps"] - st0["unrefined_steps"]).sum())
    it_hist = np.bincount(stats["total_iterations"] - 1)
    per_rank_ms = rank_all(lq_ms[0] / max(launches[1], 1))
    per_rank_iters = rank_all(iters_per_solve)
    per_rank_max_iters = rank_all(int((stats["total_iterations"] - 1).max()))

    # ---- the same step with a scheduling hint: instances started in the order of decreasing iteration counts of the previous
    # solve (cb200_lq_set_order; what a warm-started MPC loop would do).  Same results, fuller last wave; reported apart.
    launches_keep, lq_keep = list(launches), lq_ms[0]
    k.lq_set_order(np.argsort(-(stats["total_iterations"] - 1), kind="stable"))
    one_solve()
    ms_hint = timed(one_solve, max(2, args.steps // 2))
    k.lq_set_order(None)
    hint = dict(value=total_iters / (ms_hint * 1e-3), ms_per_step=ms_hint,
                what="instances started longest-first by the previous solve's iteration counts (cb200_lq_set_order); results identical")
    launches[:] = launches_keep
    lq_ms[0] = lq_keep

    # ---- dominant kernel of the step: k_lq_step.  Algorithmic bytes = SURVEY.md section 8(d): per instance B_newton = B_res +
    # B_cone + n_trials (B_asm + B_factor) + (1 + n_refine) (B_solve + B_spmv), with the factorisation / solve counts the
    # kernel actually performed.
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
    # ---- the KKT-solve unit: assemble + LDL^T + one reduced solve with recovery (k_kkt_factor_solve), roofline vs HBM
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

    def traffic_of(fname, key):
        path = os.path.join(ROOT, "profiles", fname)
        try:
            return json.load(open(path)).get(key)
        except (OSError, ValueError):
            return None
    # measured DRAM bytes of one Newton iteration of one instance (ncu capture of a launch in which every instance ran exactly
    # one iteration), scaled to the Newton iterations an average launch of the timed region carried
    per_it = traffic_of("lq_step_traffic.json", "dram_bytes_per_instance_iteration")
    traffic = per_it * args.steps * iters_per_solve / max(n_lq_launches, 1) if per_it else None
    per_inst = traffic_of("kkt_factor_solve_traffic.json", "dram_bytes_per_instance")
    kkt_traffic = per_inst * B if per_inst else None
    achieved = newton_bytes / max(n_lq_launches, 1) / (ms_lq_launch * 1e-3) / 1e9
    roofline = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic,
                    kernel="k_lq_step", launches_in_timed_region=n_lq_launches, ms_per_launch=ms_lq_launch,
                    ms_per_launch_by_rank=per_rank_ms, newton_iterations_by_rank=per_rank_iters,
                    longest_instance_by_rank=per_rank_max_iters,
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

    # ---- e2e: complete solve! of every instance through the C ABI from pinned HOST buffers.  Per step and instance: H2D of
    # the problem data (W, G, C values, q, g0, h0 -- what evaluate! produces for this family) and of the guess, initialize!,
    # solve! on the device, D2H of the solution point and of the statistics.  The batch is split over a few handles (one
    # stream each) so that the copies of one chunk overlap the kernels of the others.
    lib, A = k.lib, _lib.A
    nchunks = max(1, min(args.e2e_chunks, B // 150 if B >= 300 else 1))       # chunks stay above one CTA per SM
    bounds = [shard_range(B, c, nchunks) for c in range(nchunks)]
    chunk_handles = [BatchKKT(Ps[0], batch=e - b, device=local) for b, e in bounds]
    chunk_streams = [torch.cuda.ExternalStream(kc.lib.cb200_stream(kc.h), device=dev) for kc in chunk_handles]
    data_names = (("W_VALUES", "W_val"), ("G_VALUES", "G_val"), ("C_VALUES", "C_val"), ("LQ_Q", "q"), ("LQ_G0", "g0"), ("LQ_H0", "h0"))
    host_in = {}
    for nm, attr in data_names:
        a = np.stack([np.asarray(getattr(P, attr), dtype=np.float64) for P in plist])
        t = torch.empty(a.shape, dtype=torch.float64).pin_memory()
        t.numpy()[...] = a
        host_in[nm] = t
    host_x0 = torch.empty(X0.shape, dtype=torch.float64).pin_memory()
    host_x0.numpy()[...] = X0
    host_w = torch.empty((B, info["total"]), dtype=torch.float64).pin_memory()
    stats_out = torch.empty((B, _lib.I_COUNT), dtype=torch.int32).pin_memory()
    h2d = sum(t.numel() * 8 for t in host_in.values()) + host_x0.numel() * 8
    d2h = host_w.numel() * 8 + stats_out.numel() * 4

    def at(t, b0, ctype):
        return _lib.C.cast(t.data_ptr() + b0 * t.shape[1] * t.element_size(), ctype)

    e2e_ev = [torch.cuda.Event(enable_timing=True) for _ in range(nchunks + 1)]

    def e2e_step():
        e2e_ev[0].record(chunk_streams[0])
        for c, ((b0, e0_), kc) in enumerate(zip(bounds, chunk_handles)):
            n_, hc = e0_ - b0, kc.h
            for nm, t in host_in.items():
                lib.cb200_set_array(hc, A[nm], at(t, b0, _lib.c_dp), 0, n_)
            lib.cb200_initialize(hc, at(host_x0, b0, _lib.c_dp), 0, n_)
            lib.cb200_lq_begin(hc, 0)
            lib.cb200_lq_step(hc, args.max_newton)                  # one launch: every instance iterates to convergence
            lib.cb200_get_array_async(hc, A["POINT"], at(host_w, b0, _lib.c_dp), 0, n_)
            lib.cb200_get_stats_async(hc, at(stats_out, b0, _lib.c_ip), 0, n_)
            e2e_ev[c + 1].record(chunk_streams[c])
        for kc in chunk_handles:
            lib.cb200_synchronize(kc.h)
        return max(e2e_ev[0].elapsed_time(e) for e in e2e_ev[1:])

    for _ in range(2):
        e2e_step()
    e2e_not_converged = rank_sum(int((stats_out[:, _lib.I["converged"]] != 1).sum()))     # (reported, never asserted: ranks stay in step)
    e2e_iters = rank_sum(int((stats_out[:, _lib.I["total_iterations"]] - 1).sum()))
    barrier()
    reps = max(3, args.steps)
    ms_e2e = rank_max(sum(e2e_step() for _ in range(reps)) / reps)
    barrier()
    e2e = dict(value=e2e_iters / (ms_e2e * 1e-3), unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=ms_e2e,
               newton_iterations_per_step=e2e_iters, handles_per_gpu=nchunks, instances_not_converged=e2e_not_converged,
               what="complete solve! of every instance through the C ABI from pinned host buffers, per step: H2D of the problem "
                    "data (W/G/C values, q, g0, h0) and of the guesses, initialize!, solve! (k_lq_begin + one k_lq_step launch "
                    "per handle), D2H of the solution points and statistics; the batch is split over handles_per_gpu handles "
                    "(one stream each) so copies overlap kernels; device time from the first H2D to the last D2H, max over ranks")
    for kc in chunk_handles:
        kc.close()

    # ---- BASELINE.json's literal configurations, at every N: configs[4] = 8 instances per GPU (64 per box at 8 GPUs), and
    # (rank 0) configs[3] = ONE instance on one B200
    literal = {}

    def small_batch(bb, first_seed, make=lqc.cfg3, distinct=None):
        nd = min(bb, distinct or bb)
        Pd_ = [make(first_seed + i) for i in range(nd)]
        Pl = [Pd_[i % nd] for i in range(bb)]
        kk = BatchKKT(Pl[0], batch=bb, device=local)
        kk.load_lq(Pl)
        x0 = np.stack([P.x0 for P in Pl])
        st2 = torch.cuda.ExternalStream(kk.lib.cb200_stream(kk.h), device=dev)

        def solve2():
            kk.initialize(x0)
            kk.lq_begin()
            kk.lq_solve(max_steps=args.max_newton, check_every=args.check_every)
        for _ in range(2):
            solve2()
        its = int((kk.stats()["total_iterations"] - 1).sum())
        ms = timed(solve2, 3, st2)
        kk.lq_begin()
        kk.lq_step(4)
        kk.set_scalars(eps_p=1e-7, eps_d=1e-7)
        kk.kkt_factor_solve(1)
        ms_unit = timed(lambda: kk.kkt_factor_solve(1), 10, st2)
        small_batch.paths = kk.paths()
        small_batch.info = kk.info()
        kk.close()
        return its, ms, ms_unit

    if not args.no_literal:
        its8, ms8, unit8 = small_batch(8, 8 * rank)                    # seeds 0..63 over 8 ranks: BASELINE configs[4]
        tot8 = rank_sum(its8)
        literal["cfg4_8_instances_per_gpu"] = dict(instances_per_gpu=8, instances=8 * world, newton_it_per_s=tot8 / (ms8 * 1e-3),
                                                   ms_per_solve=ms8, newton_iterations=tot8, kkt_factor_solve_ms=unit8,
                                                   what="BASELINE.json configs[4]: seeds 8*rank .. 8*rank+7 per GPU, max over ranks")
        # the quadruped shape SURVEY.md cites (test/examples/quadruped_gait.jl: N_sym = 6954, 76-variable stages; here N = 6987,
        # 75-variable stages): the leaves-first shared-memory plan keeps three resident CTAs per SM; one wave of 444
        itsq, msq, unitq = small_batch(444, 8 * rank, make=lqc.quadruped_shape, distinct=8)
        totq = rank_sum(itsq)
        literal["quadruped_shape_batch"] = dict(instances_per_gpu=444, N=int(small_batch.info["N"]), nnz_L=int(small_batch.info["nnzL"]),
                                                newton_it_per_s=totq / (msq * 1e-3), ms_per_solve=msq, newton_iterations=totq,
                                                kkt_factor_solve_ms=unitq, code_paths=small_batch.paths,
                                                what="complete solve! of 444 quadruped-shaped instances per GPU (8 distinct seeds); "
                                                     "kkt_factor_solve_ms = one launch of the KKT-solve unit over the batch")
        # configs[3]: one instance; every rank runs it (the timing helper contains collectives), rank 0's number is reported
        its1, ms1, unit1 = small_batch(1, 0)
        literal["cfg3_single_instance"] = dict(batch=1, newton_it_per_s=its1 / (ms1 * 1e-3), ms_per_solve=ms1,
                                               newton_iterations=its1, kkt_factor_solve_ms=unit1,
                                               what="BASELINE.json configs[3]: one instance on one B200 (one CTA: latency-bound)")

    # ---- differentiate!: solution sensitivities of a few instances for many parameters -- one factorisation per
    # instance, one CTA per (instance, parameter) column; through the C ABI with host buffers (copies inside the timing)
    diff = None
    if not args.no_literal:
        nb, npar = 8, 64
        Pd = [lqc.cfg3(8 * rank + i) for i in range(nb)]
        kd = BatchKKT(Pd[0], batch=nb, device=local)
        kd.load_lq(Pd)
        kd.initialize(np.stack([P.x0 for P in Pd]))
        kd.lq_begin()
        kd.lq_solve(max_steps=args.max_newton, check_every=args.check_every)       # sensitivities are taken at the solution
        Hd = torch.empty((nb * npar, info["total"]), dtype=torch.float64).pin_memory()      # one right-hand side per row
        Hd.numpy()[...] = np.random.default_rng(rank).standard_normal(Hd.shape)
        Sd = torch.empty_like(Hd).pin_memory()
        std = torch.cuda.ExternalStream(kd.lib.cb200_stream(kd.h), device=dev)

        def diff_call():
            kd.b.check(kd.lib.cb200_differentiate(kd.h, npar, _lib.C.cast(Hd.data_ptr(), _lib.c_dp), _lib.C.cast(Sd.data_ptr(), _lib.c_dp)))
        diff_call()
        ms_d = timed(diff_call, 5, std)
        cols = nb * npar * world
        diff = dict(instances_per_gpu=nb, parameters=npar, ms_per_call=ms_d, sensitivity_columns_per_second=cols / (ms_d * 1e-3),
                    algorithmic_bytes_per_gpu=nb * (b_asm + b_factor) + nb * npar * b_solve,
                    achieved_gbs_incl_copies=(nb * (b_asm + b_factor) + nb * npar * b_solve) / (ms_d * 1e-3) / 1e9,
                    h2d_bytes=Hd.numel() * 8, d2h_bytes=Sd.numel() * 8,
                    what="cb200_differentiate on 8 converged cfg3 instances per GPU x 64 parameter columns (random dR/dtheta): H2D, "
                         "assemble + factor once per instance, 512 reduced solves with recovery on 512 CTAs, D2H")
        kd.close()

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            per = 6
            pool = OraclePool(1, per)
            pool.step(1)                                   # warm (also prepares: generation + symbolic, untimed)
            it_r, w_r, _ = pool.step(1)
            it_d, w_d, _ = pool.step(0)
            cpu = dict(value=it_r / w_r, unit=UNIT, cores=1, kind="port",
                       sample=f"{per} complete solve! runs of cfg3 (seeds 0..{per - 1}), {it_r} Newton iterations, {w_r:.1f} s on 1 core, "
                              f"reference schedule (>=3 factorisations per step); set-up outside the timed region",
                       deduplicated_schedule_value=it_d / w_d, kkt_factor_solve_ms=oracle_kkt_solve_ms())
        cfg = workload_config(B, distinct, args.check_every)
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                    data="synthetic", config=cfg,
                    symbolic=dict(nnz_K_upper=info["nnzK"], nnz_L=info["nnzL"], supernodes=info["supernodes"], levels=info["levels"],
                                  ordering=os.environ.get("CB200_ORDERING", "mindeg")),
                    newton_iterations_per_step=total_iters, with_schedule_hint=hint,
                    iterations_histogram={int(i): int(c) for i, c in enumerate(it_hist) if c},
                    e2e=e2e, gpu_launches=gpu_launches, roofline=roofline, cpu_baseline=cpu, clocks=clocks,
                    converged=conv, fallbacks_per_solve=fallbacks_per_solve, unrefined_steps=unrefined, differentiate=diff, **literal)
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
    ap.add_argument("--distinct", type=int, default=166,
                    help="distinct seeds per GPU (seeds rank*distinct + i), tiled over the batch; 0: all distinct.  Default 166: the 8 ranks "
                         "of a box then use seeds 0..1327, on all of which the reference algorithm converges (checked with the oracle and "
                         "on the GPU); a few seeds beyond (e.g. in 1332..2663) do not, and one such instance would run to the iteration "
                         "limit and dominate the step")
    ap.add_argument("--max-newton", type=int, default=400)
    ap.add_argument("--check-every", type=int, default=400,
                    help="Newton iterations per k_lq_step launch / convergence check (default: run every instance to "
                         "convergence inside one launch)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-literal", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="handles (streams) the e2e step pipelines the batch over")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
