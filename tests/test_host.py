"""CPU-side tests: the C-ABI library loads and exports every symbol include/lpc.h declares (no compute without a
GPU), the host mirror's table ordering, and the synthetic workload generators."""
import ctypes
import hashlib
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = []
    for h in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if h.endswith(".h"):
            src = open(os.path.join(ROOT, "include", h)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names += re.findall(r"\b(lpc_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    import lala_pc_b200 as L
    names = declared_functions()
    assert len(names) >= 40
    lib = ctypes.CDLL(L.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/*.h but not exported by liblpc.so"
        assert n in L.SIGNATURES, f"{n} has no ctypes signature in the host binding"
    assert b"sm_100a" in L.lib.lpc_version()


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to create handles (there is no CPU path)."""
    import lala_pc_b200 as L
    if L.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(L.LpcError):
        L.device_init(0)
    with pytest.raises(L.LpcError):
        L.Store(8)
    with pytest.raises(L.LpcError):
        L.Table(np.array([[2, 0, 1, 2]], dtype=np.int32), 3)


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "lala-pc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")) and "facade/tests" not in dirpath:
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower() or f == "__init__.py" and "oracle" not in src, (dirpath, f)


def test_sort_records_is_the_reference_order():
    """pir.hpp:343-347: stable sort by (op, y, x, z)."""
    import lala_pc_b200 as L
    rng = np.random.default_rng(0)
    r = rng.integers(0, 5, (500, 4)).astype(np.int32)
    got = L.sort_records(r)
    want = sorted(map(tuple, r.tolist()), key=lambda t: (t[0], t[2], t[1], t[3]))
    assert [tuple(x) for x in got.tolist()] == want


def test_workloads_are_deterministic_and_satisfiable():
    from lala_pc_b200 import workloads as W
    from oracle import oracle as O
    a, b = W.config1(), W.config1()
    assert np.array_equal(a.records, b.records) and np.array_equal(a.store, b.store)
    assert a.records.shape == (50_000, 4) and a.store.shape == (10_000, 2)
    counts = np.bincount(a.records[:, 0], minlength=49)
    assert counts[W.ADD] == 25_000 and counts[W.MUL] == 12_500 and counts[W.LEQ] == 12_500
    assert W.check_solution(a.records, a.solution.astype(np.int64))
    assert np.array_equal(a.records, W.sort_records(a.records))
    digest = hashlib.sha256(a.records.tobytes() + a.store.tobytes()).hexdigest()
    gold = open(os.path.join(ROOT, "tests", "golden", "workloads.sha256")).read().split()
    assert digest == gold[0], "config 1 generator output changed: regenerate tests/golden/workloads.sha256"
    # the planted solution survives propagation, the failing twin does not
    s, st = O.pir_fixpoint(a.store, a.records)
    assert not st.is_bot and st.sweeps > 3
    assert ((s[:, 0] <= a.solution) & (a.solution <= s[:, 1])).all()
    s2, st2 = O.pir_fixpoint(a.failing_twin().store, a.records)
    assert st2.is_bot


def test_eps_split_shapes():
    from lala_pc_b200 import workloads as W
    from oracle import oracle as O
    net = W.config4_base()
    assert net.records.shape == (10_000, 4) and net.store.shape == (2_000, 2)
    root, st = O.pir_fixpoint(net.store, net.records)
    dec, obj = W.eps_decisions(net.records, root)
    assert len(set(dec)) == 16 and obj not in dec
    stores = W.eps_stores(root, dec, 0, 8)
    for k in range(8):
        for j, v in enumerate(dec):
            lb, ub = root[v]
            mid = lb + ((ub - lb) >> 1)
            want = (mid + 1, ub) if (k >> j) & 1 else (lb, mid)
            assert tuple(stores[k, v]) == want
    untouched = np.setdiff1d(np.arange(2000), dec)
    assert np.array_equal(stores[3][untouched], root[untouched])
