"""The drop-in boundary: libcalipso_b200.so loads without a GPU and exports every symbol include/calipso_b200.h declares
(no compute calls here); the Python binding table covers the same set; the test-only emulation exports it too."""
import ctypes
import os
import re

import backends
from calipso_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "calipso_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cb200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    names = declared_symbols()
    for must in ("cb200_create", "cb200_ldl_create", "cb200_residual", "cb200_search_direction", "cb200_cone",
                 "cb200_cone_search", "cb200_apply_step", "cb200_ldl_factorize", "cb200_ldl_inertia", "cb200_ldl_solve",
                 "cb200_ldl_linear_solve", "cb200_allreduce_counts"):
        assert must in names


def test_cuda_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(build.build())
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_table_matches_header():
    assert sorted(_lib.SYMBOLS) == declared_symbols()


def test_emulation_exports_the_same_boundary():
    lib = backends.binding("emul").lib
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback_without_a_device():
    """cb200_create must fail loudly (not fall back) when there is no CUDA device."""
    b = _lib.Binding(build.build())
    if b.lib.cb200_device_count() > 0:
        return
    from calipso_b200 import lqc
    from calipso_b200.solver import BatchKKT
    import pytest
    with pytest.raises(Exception, match="no CUDA device"):
        BatchKKT(lqc.tiny(), binding=b)


def test_async_readback_equals_blocking_readback():
    """cb200_get_array_async / cb200_get_stats_async + cb200_synchronize return what the blocking calls return
    (emulation backend here; the GPU path is exercised by bench.py's pipelined e2e step)."""
    import numpy as np
    from calipso_b200 import lqc
    from calipso_b200.solver import BatchKKT
    b = backends.binding("emul")
    P = lqc.tiny()
    k = BatchKKT(P, batch=2, binding=b)
    k.load_lq([P, lqc.tiny(1)])
    k.initialize(np.stack([P.x0, lqc.tiny(1).x0]))
    k.lq_begin()
    k.lq_step(2)
    ref = k.get("POINT")
    out = np.zeros_like(ref)
    assert b.lib.cb200_get_array_async(k.h, _lib.A["POINT"], _lib.dp(out), 0, 2) == 0
    st = np.zeros((2, _lib.I_COUNT), dtype=np.int32)
    assert b.lib.cb200_get_stats_async(k.h, _lib.ip(st), 0, 2) == 0
    assert b.lib.cb200_synchronize(k.h) == 0
    assert np.array_equal(out, ref)
    assert np.array_equal(st[:, _lib.I["total_iterations"]], k.stats()["total_iterations"])
    assert b.lib.cb200_values_changed(k.h) == 0


import pytest  # noqa: E402


@pytest.mark.parametrize("backend", backends.BACKENDS)
def test_initialize_writes_the_primal_block_only(backend):
    """cb200_initialize = initialize!(solver, guess), initialize.jl:9-13: solution.variables .= guess, nothing else."""
    import numpy as np
    from calipso_b200 import lqc
    from calipso_b200.solver import BatchKKT
    P = lqc.tiny()
    k = BatchKKT(P, batch=4, binding=backends.binding(backend))
    rng = np.random.default_rng(0)
    w0 = rng.standard_normal((4, k.total))
    k.set("POINT", w0)
    g = rng.standard_normal((4, k.n))
    k.initialize(g)
    w = k.get("POINT")
    assert np.array_equal(w[:, :k.n], g) and np.array_equal(w[:, k.n:], w0[:, k.n:])
    g2 = rng.standard_normal((2, k.n))                       # a sub-range of the batch through the C ABI
    k.b.check(k.lib.cb200_initialize(k.h, _lib.dp(_lib.f64(g2)), 1, 2))
    w2 = k.get("POINT")
    assert np.array_equal(w2[1:3, :k.n], g2) and np.array_equal(w2[[0, 3]], w[[0, 3]])
    with pytest.raises(_lib.CalipsoB200Error):
        k.b.check(k.lib.cb200_initialize(k.h, _lib.dp(_lib.f64(g2)), 3, 2))
