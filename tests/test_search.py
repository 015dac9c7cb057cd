"""In-kernel search (include/lpc.h: lpc_batch_search): the CPU restatement (oracle.pir_search) on
problems with known solution counts, and the device search against it, subproblem by subproblem."""
import itertools

import numpy as np
import pytest

from oracle import oracle as O

ADD, MUL, EQ, LEQ = 2, 4, 46, 48


def all_different(n, lo, hi):
    """n variables in [lo, hi], pairwise different: EQ(X = ZERO, Y = a, Z = b) for every pair (pir_test.cpp:550-560)."""
    zero = n
    recs = [[EQ, zero, a, b] for a, b in itertools.combinations(range(n), 2)]
    store = [[lo, hi]] * n + [[0, 0]]
    if (n + 1) % 2:
        store.append([0, 0])        # batched stores need an even number of variables
    return np.array(recs, dtype=np.int32), np.array(store, dtype=np.int32)


def test_oracle_search_known_counts():
    # x + y = 3 over [0,3]^2: 4 solutions in a complete binary tree of 7 nodes
    recs = np.array([[ADD, 2, 0, 1]], dtype=np.int32)
    out = O.pir_search(np.array([[[0, 3], [0, 3], [3, 3]]], dtype=np.int32), recs, [0, 1, 2])
    assert out[0].tolist() == [4, 7, 0, 2**31 - 1, 0, 0]
    # permutations of 1..n
    for n, want in ((3, 6), (4, 24), (5, 120)):
        recs, store = all_different(n, 1, n)
        out = O.pir_search(store[None], recs, list(range(n)), objective_var=0)
        assert out[0, 0] == want and out[0, 3] == 1 and out[0, 4] == 0 and out[0, 5] == 0
        assert out[0, 1] == 2 * (out[0, 0] + out[0, 2]) - 1          # a full binary tree: nodes = 2 * leaves - 1
    # x * y = 12 over [1,12]^2: the divisor pairs
    recs = np.array([[MUL, 2, 0, 1]], dtype=np.int32)
    out = O.pir_search(np.array([[[1, 12], [1, 12], [12, 12], [0, 0]]], dtype=np.int32), recs, [0, 1])
    assert out[0, 0] == 6
    # limits: a node budget abandons the subproblem
    recs, store = all_different(5, 1, 5)
    out = O.pir_search(store[None], recs, list(range(5)), max_nodes=10)
    assert out[0, 4] == 1 and 10 <= out[0, 1] < 239
    out = O.pir_search(store[None], recs, list(range(5)), max_depth=2)
    assert out[0, 4] == 1


def random_model(rng, nvars):
    recs = []
    for _ in range(int(rng.integers(2, 9))):
        op = int(rng.choice([ADD, ADD, MUL, LEQ, EQ]))
        x, y, z = (int(v) for v in rng.permutation(nvars)[:3])
        recs.append([op, x, y, z])
    a = rng.integers(-3, 6, (nvars, 2))
    store = np.stack([a.min(1), a.max(1)], axis=1).astype(np.int32)
    for r in recs:
        if r[0] in (EQ, LEQ):
            store[r[1]] = (0, 1)
    return np.array(recs, dtype=np.int32), store


@pytest.fixture(scope="module")
def L():
    import lala_pc_b200 as L
    L.device_init(0)
    return L


@pytest.mark.gpu
def test_device_search_matches_oracle_small_models(L):
    rng = np.random.default_rng(17)
    total_solutions = 0
    for trial in range(40):
        nvars = 6
        recs, store = random_model(rng, nvars)
        # 16 roots: the model with its first two variables pre-split, like an EPS decomposition
        roots = np.repeat(store[None], 16, axis=0).copy()
        for k in range(16):
            roots[k, 0] = (store[0, 0] + (k % 4) % (store[0, 1] - store[0, 0] + 1),) * 2
            roots[k, 1] = (store[1, 0] + (k // 4) % (store[1, 1] - store[1, 0] + 1),) * 2
        bv = list(range(nvars))
        want = O.pir_search(roots, recs, bv, objective_var=2)
        t = L.Table(recs, nvars)
        b = L.Batch(t, 16)
        b.write(roots)
        r, got = b.search(bv, objective_var=2, change_driven=bool(trial & 1))
        assert np.array_equal(got, want), (trial, got.tolist(), want.tolist())
        assert r.n_solutions == want[:, 0].sum() and r.n_nodes == want[:, 1].sum() and r.n_fails == want[:, 2].sum()
        assert r.best_bound == want[:, 3].min()
        assert np.array_equal(b.read(), roots)          # the roots are left untouched
        total_solutions += int(r.n_solutions)
        b.close()
    assert total_solutions > 100


@pytest.mark.gpu
def test_device_search_permutations_and_limits(L):
    recs, store = all_different(6, 1, 6)
    nvars = store.shape[0]
    t = L.Table(recs, nvars)
    # EPS: 64 subproblems = the halves of three decision variables ... here simply x0, x1 fixed to each of 36 values
    roots = []
    for a in range(1, 7):
        for c in range(1, 7):
            s = store.copy(); s[0] = (a, a); s[1] = (c, c); roots.append(s)
    roots = np.array(roots, dtype=np.int32)
    b = L.Batch(t, len(roots))
    b.write(roots)
    bv = list(range(6))
    r, got = b.search(bv, objective_var=5)
    assert r.n_solutions == 720 and r.n_incomplete == 0 and r.n_unknown_leaves == 0 and r.best_bound == 1
    want = O.pir_search(roots, recs, bv, objective_var=5)
    assert np.array_equal(got, want)
    # budgets: same abandoned subproblems, same counts up to the point of abandonment
    for kw in (dict(max_nodes=5), dict(max_depth=2), dict(max_nodes=9, max_depth=3)):
        r, got = b.search(bv, objective_var=5, **kw)
        want = O.pir_search(roots, recs, bv, objective_var=5, **kw)
        assert np.array_equal(got, want), kw
        assert r.n_incomplete == want[:, 4].sum() > 0
    b.close()


@pytest.mark.gpu
def test_device_search_config4_model_with_node_budget(L):
    """The BASELINE.json config-4 model: 256 EPS subproblems searched with a budget of 40 nodes each; every per-store
    record (solutions, nodes, failures, best, abandoned) equals the CPU restatement's."""
    from lala_pc_b200 import workloads as W
    net = W.config4_base()
    root, st = O.pir_fixpoint(net.store, net.records)
    dec, obj = W.eps_decisions(net.records, root, n=8)
    stores = W.eps_stores(root, dec, 0, 256)
    width = root[:, 1].astype(np.int64) - root[:, 0]
    bv = [int(v) for v in np.argsort(-width, kind="stable")[:48]]
    want = O.pir_search(stores, net.records, bv, objective_var=obj, max_nodes=40, max_depth=32, threads=8)
    t = L.Table(net.records, net.nvars)
    b = L.Batch(t, 256)
    b.write(stores)
    r, got = b.search(bv, objective_var=obj, max_nodes=40, max_depth=32, change_driven=False)
    assert np.array_equal(got, want), np.flatnonzero((got != want).any(1))[:8].tolist()
    assert r.n_nodes == want[:, 1].sum() and r.n_nodes > 256
    # the roots are the root fixpoint except on the decision variables: with that promise the root nodes are incremental too
    # change-driven nodes: every fixpoint starts from the propagators of the branched variable; with the promise that the
    # roots are the root fixpoint except on the decision variables the root nodes are incremental too. Same records.
    r1, got1 = b.search(bv, objective_var=obj, max_nodes=40, max_depth=32, change_driven=True)
    assert np.array_equal(got1, want) and r1.deductions < r.deductions
    b.set_seeds(dec)
    r2, got2 = b.search(bv, objective_var=obj, max_nodes=40, max_depth=32, change_driven=True)
    assert np.array_equal(got2, want) and r2.deductions < r1.deductions
    b.close()
