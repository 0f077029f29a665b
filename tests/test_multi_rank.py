"""N > 1 host logic on CPU: two gloo ranks shard a batch of instances (no data-path collective), all-reduce the four
convergence counters, and together reproduce the single-process result."""
import os
import socket
import sys

import numpy as np
import pytest

from calipso_b200.sharding import shard_range

HERE = os.path.dirname(os.path.abspath(__file__))


def test_shard_range_partitions():
    for total in (0, 1, 5, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, total, out):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    import backends
    from calipso_b200 import lqc
    from calipso_b200.sharding import solve_sharded
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        Ps = [lqc.tiny(i) for i in range(total)]
        k, glob, (b, e) = solve_sharded(Ps, rank, world, binding=backends.binding("emul"), max_steps=300, check_every=3)
        np.savez(out, begin=b, end=e, W=k.get("POINT") if k is not None else np.zeros((0, 1)),
                 it=k.stats()["total_iterations"] if k is not None else np.zeros(0), **{f"g_{kk}": v for kk, v in glob.items()})
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_gloo_ranks_match_single_process(tmp_path):
    import torch.multiprocessing as mp
    import backends
    from calipso_b200 import lqc
    from calipso_b200.solver import BatchKKT
    total, world = 5, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    outs = [str(tmp_path / f"rank{r}.npz") for r in range(world)]
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, outs[r])) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    # single-process reference
    Ps = [lqc.tiny(i) for i in range(total)]
    k = BatchKKT(Ps[0], batch=total, binding=backends.binding("emul"))
    k.load_lq(Ps)
    k.initialize(np.stack([P.x0 for P in Ps]))
    k.lq_begin()
    r = k.lq_solve(max_steps=300, check_every=3)
    W, it = k.get("POINT"), k.stats()["total_iterations"]
    seen = 0
    for o in outs:
        d = np.load(o)
        b, e = int(d["begin"]), int(d["end"])
        assert np.array_equal(d["W"], W[b:e])              # same code, same inputs: bit-identical
        assert np.array_equal(d["it"], it[b:e])
        assert int(d["g_converged"]) == r["converged"] == total and int(d["g_running"]) == 0
        seen += e - b
    assert seen == total
