"""Parity tests (-m gpu) of the batched mode at BASELINE.json's full config-4 size and of the EPS-native call
(include/lpc.h: lpc_eps_*), through the C-ABI, against the CPU oracle: flags, every non-failed store, the reduction
record. Reference semantics: pir.hpp:721-817 (deduce), 873-884 (is_extractable), 417-438 (ask)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ADD, MUL, MIN, MAX, TDIV, FDIV, CDIV, EDIV, EQ, LEQ = 2, 4, 6, 7, 25, 27, 29, 31, 46, 48


@pytest.fixture(scope="module")
def L():
    import lala_pc_b200 as L
    L.device_init(0)
    return L


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


@pytest.fixture(scope="module")
def W():
    from lala_pc_b200 import workloads
    return workloads


@pytest.fixture(scope="module")
def c4(O, W):
    """Config 4 at its stated size: 65,536 EPS subproblems of the 2k-variable / 10k-propagator model, and what the CPU
    checker makes of every one of them (one store per host thread, Gauss-Seidel each)."""
    net = W.config4_base()
    root, st = O.pir_fixpoint(net.store, net.records)
    assert not st.is_bot
    dec, obj = W.eps_decisions(net.records, root, n=24)   # bench.py's decision list: the first 16 of the 24 widest
    dec = dec[:16]
    n = 65536
    stores = W.eps_stores(root, dec, 0, n)
    want, wflags, _, _, _ = O.pir_batch_fixpoint(stores, net.records, threads=os.cpu_count() or 8)
    del stores
    ok = (wflags & 1) == 0
    alive = np.flatnonzero(ok)
    return dict(net=net, root=root, dec=dec, obj=obj, n=n, wflags=wflags, alive=alive, want_alive=want[alive].copy(),
                n_bot=int((~ok).sum()), n_sol=int(((wflags & 2) != 0).sum()), best=int(want[alive][:, obj, 0].min()))


def check_record(res, c4):
    assert res.n_bot == c4["n_bot"] and res.n_solution == c4["n_sol"]
    assert res.n_unknown == c4["n"] - c4["n_bot"] - c4["n_sol"]
    assert res.best_bound == c4["best"]


@pytest.mark.parametrize("mode_name", ["sweep", "auto"])
def test_config4_full_size_resident(L, c4, mode_name):
    """lpc_batch_init_split + lpc_batch_fixpoint on all 65,536 stores (the grouped kernel over the packed table; in AUTO
    mode with the propagators entailed on the root dropped): every flag, every non-failed store, the reduction record."""
    net = c4["net"]
    t = L.Table(net.records, net.nvars)
    b = L.Batch(t, c4["n"])
    b.init_split(c4["root"], c4["dec"], 0)
    res = b.fixpoint(objective_var=c4["obj"], mode=L.MODE_SWEEP if mode_name == "sweep" else L.MODE_AUTO)
    flags = b.flags()
    assert np.array_equal(flags, c4["wflags"])
    got = b.read()
    assert np.array_equal(got[c4["alive"]], c4["want_alive"])
    check_record(res, c4)
    # deduce() evaluations actually executed: a failed store's last sweep is cut short (and a run of odd length evaluates
    # its padding record twice), so the count is at most sweeps x propagators, give or take the padding
    assert 0 < res.deductions <= res.sweeps_total * (len(net.records) + 16)
    if mode_name == "auto":
        assert res.deductions < res.sweeps_total * len(net.records) * 0.7   # entailed propagators were dropped
        # a second fixpoint on the same images (no longer the root of the split outside the decision variables: the
        # first-sweep shortcut must be off) finds the surviving stores at rest (failed images are not written back: they
        # fail again)
        res2 = b.fixpoint(objective_var=c4["obj"], mode=L.MODE_AUTO)
        assert np.array_equal(b.flags(), c4["wflags"]) and np.array_equal(b.read()[c4["alive"]], c4["want_alive"])
        assert res2.sweeps_total < res.sweeps_total
        # and a fresh split after it turns it on again
        b.init_split(c4["root"], c4["dec"], 0)
        res3 = b.fixpoint(objective_var=c4["obj"], mode=L.MODE_AUTO)
        assert np.array_equal(b.flags(), c4["wflags"]) and np.array_equal(b.read()[c4["alive"]], c4["want_alive"])
        assert abs(res3.deductions - res.deductions) < 0.02 * res.deductions
    b.close()


@pytest.mark.parametrize("mode_name", ["sweep", "auto"])
def test_config4_full_size_eps(L, c4, mode_name):
    """The EPS-native call on the same 65,536 subproblems: root + decisions in, flags + compacted survivors out."""
    net = c4["net"]
    t = L.Table(net.records, net.nvars)
    e = L.Eps(t, c4["n"], survivor_cap=8192)
    flags = np.zeros(c4["n"], dtype=np.uint8)
    surv = np.zeros((8192, net.nvars, 2), dtype=np.int32)
    idx = np.zeros(8192, dtype=np.int32)
    res, nw = e.solve_host(c4["root"], c4["dec"], first_id=0, n=c4["n"], objective_var=c4["obj"], flags=flags,
                           survivors=surv, survivor_index=idx, mode=L.MODE_SWEEP if mode_name == "sweep" else L.MODE_AUTO)
    assert np.array_equal(flags, c4["wflags"])
    check_record(res, c4)
    assert res.n_survivors == nw == len(c4["alive"])
    order = np.argsort(idx[:nw])
    assert np.array_equal(idx[:nw][order], c4["alive"])
    assert np.array_equal(surv[:nw][order], c4["want_alive"])
    if mode_name == "auto":
        assert 0 < res.n_live_records < len(net.records)
        # the root is a fixpoint: the first sweep of every subproblem ran on the records that mention a decision variable
        assert 0 < res.n_first_sweep_records < res.n_live_records // 4
    else:
        assert res.n_live_records == len(net.records) and res.n_first_sweep_records == 0
    assert 0 < res.deductions <= res.sweeps_total * (res.n_live_records + 16)
    e.close()


def test_eps_root_that_is_not_a_fixpoint(L, O, W):
    """The first-sweep shortcut needs a root that is a common fixpoint of the table; a root that is not one (here: the
    initial store of the model, never propagated) must turn it off and still give the exact fixpoints."""
    net = W.pir_network(600, 2400, seed=91, width=10)
    raw = net.store.copy()
    fix, st = O.pir_fixpoint(raw, net.records)
    assert not st.is_bot and not np.array_equal(fix, raw)
    dec, obj = W.eps_decisions(net.records, fix, n=10, min_degree=2)
    ids = np.arange(1 << len(dec), dtype=np.int64)
    for root, expect_first in ((raw, False), (fix, True)):
        want, wflags = oracle_eps(O, W, net.records, root, dec, ids)
        t = L.Table(net.records, net.nvars)
        e = L.Eps(t, len(ids), survivor_cap=len(ids))
        flags = np.zeros(len(ids), dtype=np.uint8)
        surv = np.zeros((len(ids), net.nvars, 2), dtype=np.int32)
        idx = np.zeros(len(ids), dtype=np.int32)
        res, nw = e.solve_host(root, dec, ids=ids, objective_var=obj, flags=flags, survivors=surv, survivor_index=idx)
        assert (res.n_first_sweep_records > 0) == expect_first, (expect_first, res.n_first_sweep_records)
        assert np.array_equal(flags, wflags)
        ok = (wflags & 1) == 0
        assert nw == int(ok.sum())
        assert np.array_equal(surv[:nw], want[idx[:nw]])
        e.close()
        t.close()


def oracle_eps(O, W, records, root, dec, ids):
    stores = W.eps_stores(root, dec, 0, len(ids), ids=ids)
    want, wflags, wsweeps, _, _ = O.pir_batch_fixpoint(stores, records, threads=8)
    return want, wflags


def check_eps(L, O, W, records, nvars, root, dec, ids, obj=-1, cap=None, mode=None, label=""):
    want, wflags = oracle_eps(O, W, records, root, dec, ids)
    ok = (wflags & 1) == 0
    t = L.Table(records, nvars)
    e = L.Eps(t, len(ids), survivor_cap=cap)
    kw = {} if mode is None else dict(mode=mode)
    for rep in range(2):   # the handle is reusable
        flags = np.full(len(ids), 0xff, dtype=np.uint8)
        ncap = e.survivor_cap
        surv = np.zeros((max(ncap, 1), nvars, 2), dtype=np.int32)
        idx = np.full(max(ncap, 1), -1, dtype=np.int32)
        res, nw = e.solve_host(root, dec, ids=ids, objective_var=obj, flags=flags, survivors=surv, survivor_index=idx, **kw)
        assert np.array_equal(flags, wflags), (label, rep)
        assert res.n_bot == int((~ok).sum()) and res.n_solution == int(((wflags & 2) != 0).sum()), (label, rep)
        assert res.n_survivors == int(ok.sum()), (label, rep)
        assert nw == min(int(ok.sum()), ncap), (label, rep)
        got_idx = idx[:nw]
        assert len(set(got_idx.tolist())) == nw and ok[got_idx].all(), (label, rep)
        assert np.array_equal(surv[:nw], want[got_idx]), (label, rep)
        if obj >= 0 and ok.any():
            assert res.best_bound == int(want[ok][:, obj, 0].min()), (label, rep)
    e.close()
    t.close()
    return wflags


def test_eps_scrambled_ids_and_small_capacity(L, O, W):
    """Explicit (scrambled) subproblem ids as a rank of a multi-GPU run gets them; a survivor buffer smaller than the
    number of survivors keeps that many and still reports them all; a root that is NOT a fixpoint."""
    from lala_pc_b200 import sharding
    net = W.config4_base()
    root, _ = O.pir_fixpoint(net.store, net.records)
    dec, obj = W.eps_decisions(net.records, root, n=19)
    ids = sharding.shard_ids(5, 8, 65536)[1000:1000 + 3000]
    wflags = check_eps(L, O, W, net.records, net.nvars, root, dec, ids, obj=obj, label="scrambled")
    n_alive = int(((wflags & 1) == 0).sum())
    assert n_alive > 8
    check_eps(L, O, W, net.records, net.nvars, root, dec, ids, obj=obj, cap=n_alive // 2, label="small capacity")
    check_eps(L, O, W, net.records, net.nvars, root, dec, ids, obj=obj, mode=L.MODE_SWEEP, label="no elimination")
    # the unpropagated model as root: nothing is a fixpoint yet, subproblems take many sweeps
    dec2, obj2 = W.eps_decisions(net.records, net.store, n=10)
    check_eps(L, O, W, net.records, net.nvars, net.store, dec2, np.arange(1024, dtype=np.int64), obj=obj2, label="raw root")


def test_eps_edge_cases(L, O, W):
    net = W.config4_base()
    root, _ = O.pir_fixpoint(net.store, net.records)
    dec, obj = W.eps_decisions(net.records, root, n=6)
    # no decisions: every subproblem is the root itself
    check_eps(L, O, W, net.records, net.nvars, root, [], np.zeros(5, dtype=np.int64), obj=obj, label="no decisions")
    # a single subproblem
    check_eps(L, O, W, net.records, net.nvars, root, dec, np.array([37], dtype=np.int64), label="one id")
    # a root with an empty variable: every subproblem fails before the first sweep
    bad = root.copy()
    bad[3] = (5, 4)
    wflags = check_eps(L, O, W, net.records, net.nvars, bad, dec, np.arange(64, dtype=np.int64), label="bot root")
    assert (wflags & 1).all()
    # halving a singleton decision variable with bit 1 empties it
    one = root.copy()
    one[dec[0]] = (one[dec[0], 0], one[dec[0], 0])
    check_eps(L, O, W, net.records, net.nvars, one, dec, np.arange(64, dtype=np.int64), label="singleton decision")
    # an empty table: nothing to propagate, every non-empty store is a solution
    t = L.Table(np.zeros((0, 4), dtype=np.int32), 4)
    e = L.Eps(t, 8)
    r4 = np.array([[0, 9], [0, 9], [1, 1], [2, 3]], dtype=np.int32)
    flags = np.zeros(8, dtype=np.uint8)
    res, nw = e.solve_host(r4, [0, 1], first_id=0, n=4, flags=flags[:4])
    assert res.n_solution == 4 and res.n_bot == 0 and (flags[:4] == 2).all()
    # argument checks
    with pytest.raises(L.LpcError):
        e.solve_host(r4, [0, 0], first_id=0, n=4)          # repeated decision variable
    with pytest.raises(L.LpcError):
        e.solve_host(r4, [0], first_id=0, n=9)             # more subproblems than the handle holds
    with pytest.raises(L.LpcError):
        L.Eps(L.Table(np.zeros((0, 4), dtype=np.int32), 3), 8)   # odd number of variables
    e.close()


def soup(rng, nvars, nrec, ops):
    recs = np.zeros((nrec, 4), dtype=np.int32)
    recs[:, 0] = rng.choice(ops, nrec)
    recs[:, 1:] = rng.integers(0, nvars, (nrec, 3))
    lb = rng.integers(-12, 8, nvars)
    ub = lb + rng.integers(0, 14, nvars)
    store = np.stack([lb, ub], 1).astype(np.int32)
    for i in np.flatnonzero((recs[:, 0] == EQ) | (recs[:, 0] == LEQ)):
        store[recs[i, 1]] = (max(store[recs[i, 1], 0], 0), min(store[recs[i, 1], 1], 1))
    return recs, store


def test_eps_all_operators(L, O, W):
    """Random networks over all ten operators (divisions included) and tables that are NOT sorted by operator (one mixed
    run, opcode per record), with infinite bounds in the root (the guarded `+` rule)."""
    import lala_pc_b200 as LL
    rng = np.random.default_rng(20261018)
    all_ops = [ADD, MUL, MIN, MAX, TDIV, FDIV, CDIV, EDIV, EQ, LEQ]
    for trial in range(6):
        nvars = 64
        recs, store = soup(rng, nvars, 48, all_ops if trial % 2 == 0 else [ADD, MUL, MIN, MAX, EQ, LEQ])
        if trial < 4:
            recs = LL.sort_records(recs)
        if trial % 3 == 2:
            # infinite bounds only where the rules keep them exactly infinite (`+`, min, max): a reified comparison turns
            # -inf into INT_MIN + 1, after which int32 wraps and the result depends on the schedule
            safe = np.ones(nvars, dtype=bool)
            for op, x, y, z in recs:
                if op not in (ADD, MIN, MAX):
                    safe[[x, y, z]] = False
            cand = np.flatnonzero(safe)
            if len(cand):
                store[rng.choice(cand, min(6, len(cand)), replace=False), 0] = -2**31
                store[rng.choice(cand, min(6, len(cand)), replace=False), 1] = 2**31 - 1
        dec = [int(v) for v in rng.choice(nvars, 8, replace=False)]
        ids = rng.integers(0, 256, 300).astype(np.int64)
        for mode in (L.MODE_AUTO, L.MODE_SWEEP):
            check_eps(L, O, W, recs, nvars, store, dec, ids, obj=int(rng.integers(0, nvars)), mode=mode, label=f"soup {trial}")


def test_batch_payload_matches_the_record(L, O, W):
    """The all-reduce payload of a rank: the three counters, then this rank's bound in its own slot (zeros elsewhere)."""
    import torch
    net = W.config4_base()
    root, _ = O.pir_fixpoint(net.store, net.records)
    dec, obj = W.eps_decisions(net.records, root, n=12)
    t = L.Table(net.records, net.nvars)

    class _View:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}

    for n in (4096, 64):     # grouped kernel / one store per block
        for handle in ("eps", "batch"):
            if handle == "eps":
                h = L.Eps(t, n)
                h.set_rank(2, 4)
                h.upload(root, dec, first_id=0, n=n)
                res = h.run(objective_var=obj)
            else:
                h = L.Batch(t, n)
                h.set_rank(2, 4)
                h.init_split(root, dec, 0)
                res = h.fixpoint(objective_var=obj)
            ptr, cnt = h.payload
            assert cnt == 7
            p = torch.as_tensor(_View(ptr, cnt), device="cuda").cpu().tolist()
            assert p[:3] == [res.n_solution, res.n_bot, res.n_unknown], (n, handle)
            assert p[3:] == [0, 0, res.best_bound, 0], (n, handle)
            h.close()


def test_eps_zero_copy_outputs(L, O, W):
    """Pinned (device-accessible) survivor buffers are written by the kernel itself, over the link while the other
    subproblems are still being propagated; the result is what the staged path (pageable buffers) delivers."""
    import torch
    net = W.config4_base()
    root, _ = O.pir_fixpoint(net.store, net.records)
    dec, obj = W.eps_decisions(net.records, root, n=13)
    n = 8192
    want, wflags = oracle_eps(O, W, net.records, root, dec, np.arange(n, dtype=np.int64))
    ok = (wflags & 1) == 0
    t = L.Table(net.records, net.nvars)
    e = L.Eps(t, n, survivor_cap=n)
    cap = int(ok.sum()) + 7
    surv = torch.zeros((cap, net.nvars, 2), dtype=torch.int32).pin_memory()
    idx = torch.full((cap,), -1, dtype=torch.int32).pin_memory()
    flags = torch.zeros(n, dtype=torch.uint8).pin_memory()
    for rep in range(2):
        surv.zero_()
        idx.fill_(-1)
        res, nw = e.solve_host(root, dec, first_id=0, n=n, objective_var=obj, flags=flags.data_ptr(), survivors=surv.data_ptr(),
                               survivor_index=idx.data_ptr(), max_survivors=cap)
        assert np.array_equal(flags.numpy(), wflags)
        assert nw == res.n_survivors == int(ok.sum())
        got_idx = idx.numpy()[:nw]
        assert sorted(got_idx.tolist()) == np.flatnonzero(ok).tolist()
        assert np.array_equal(surv.numpy()[:nw], want[got_idx])
        assert (idx.numpy()[nw:] == -1).all()
    e.close()
