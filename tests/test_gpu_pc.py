"""Parity tests of the PC path (-m gpu): flat propagators on the device (include/lpc_pc.h) against the tree-walking
oracle (oracle/pc_oracle.cpp) and the reference's golden vectors (tests/pc_test.cpp)."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


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


def to_tree(x):
    return tuple(to_tree(e) if isinstance(e, list) else e for e in x)


def flattenable(formulas):
    from lala_pc_b200 import pcflat
    try:
        return pcflat.flatten(formulas)
    except pcflat.Unsupported:
        return None


def check_parity(L, O, formulas, store, label=""):
    from lala_pc_b200 import pcflat
    props, terms = pcflat.flatten(formulas)
    m = O.PCModel(formulas)
    want, st = m.fixpoint(store)
    t = L.PcTable(props, terms, len(store))
    s = L.Store(values=store)
    r = t.fixpoint(s)                      # default mode: lane tiles whose operands did not move are skipped
    got = s.read()
    assert bool(r.is_bot) == bool(st.is_bot), (label, "bot flag")
    sd = L.Store(values=store)
    rd = t.fixpoint(sd, mode=L.MODE_SWEEP)  # every propagator in every sweep
    assert bool(rd.is_bot) == bool(st.is_bot), (label, "bot flag, dense")
    assert rd.deductions == rd.sweeps * len(formulas) and r.deductions <= r.sweeps * len(formulas), (label, "deductions")
    if not st.is_bot:
        assert np.array_equal(sd.read(), got), (label, "dense vs tile-skipping store")
    if not st.is_bot:
        assert np.array_equal(got, want), (label, "store", np.flatnonzero((got != want).any(1))[:5].tolist())
        n_ent, bits = m.ask_all(want, want_bits=True)
        g_ent, gbits = t.ask_all(s, want_bits=True)
        assert np.array_equal(gbits, bits) and g_ent == n_ent, (label, "ask bits")
        assert bool(r.has_changed) == bool(st.has_changed), (label, "has_changed")
    return got, r, st


def test_golden_vectors(L, O):
    """Every pc_test.cpp golden whose tree has a flat kind, through the device path."""
    kats = load_golden("pc_kat.json")["props"]
    ran = 0
    for k in kats:
        formulas = [to_tree(p) for p in k["props"]]
        if flattenable(formulas) is None:
            continue
        ran += 1
        store = np.array(k["store"], dtype=np.int32)
        got, r, st = check_parity(L, O, formulas, store, k["name"])
        if k["bot"]:
            assert r.is_bot, k["name"]
            continue
        after = np.array(k["after"], dtype=np.int32)
        assert np.array_equal(got[:len(after)], after), (k["name"], got.tolist())
    assert ran == len(kats) >= 96, ran   # flat kinds + tree propagators: every reference golden runs on the device


def test_deduce_one_step_by_step(L, O):
    """PC::deduce(i) one propagator at a time in index order == one Gauss-Seidel sweep of the tree walker."""
    from lala_pc_b200 import pcflat
    for k in load_golden("pc_kat.json")["props"]:
        formulas = [to_tree(p) for p in k["props"]]
        flat = flattenable(formulas)
        if flat is None:
            continue
        m = O.PCModel(formulas)
        store = np.array(k["store"], dtype=np.int32)
        t = L.PcTable(flat[0], flat[1], len(store))
        s = L.Store(values=store)
        bot = False
        for sweep in range(3):
            for i in range(len(formulas)):
                store, changed, bot = m.deduce(i, store, bot)
                assert t.deduce(s, i) == changed, (k["name"], sweep, i)
                if not bot:
                    assert np.array_equal(s.read(), store), (k["name"], sweep, i)


@pytest.mark.parametrize("cfg", ["config3", "config5"])
def test_config_shapes_parity_small(L, O, W, cfg):
    net = getattr(W, cfg)(0.05)
    formulas = net.formulas()
    got, r, st = check_parity(L, O, formulas, net.store, cfg)
    assert not st.is_bot and st.sweeps >= 3
    assert ((got[:, 0] <= net.solution) & (net.solution <= got[:, 1])).all()
    # failing twin: push one variable of the first propagator far above its planted value
    twin = net.store.copy()
    v = int(net.terms[net.props[0, 1], 1])
    twin[v] = (net.solution[v] + 1000, net.solution[v] + 1001)
    _, r2, st2 = check_parity(L, O, formulas, twin, cfg + " twin")
    assert st2.is_bot and r2.is_bot


def test_config3_full_size(L, O, W):
    """BASELINE.json config 3 at full size (200k vars, 1M terms): bit-exact against the tree walker; idempotent."""
    net = W.config3()
    assert net.meta["nterms"] >= 1_000_000
    m = O.PCModel(net.formulas())
    want, st = m.fixpoint(net.store)
    t = L.PcTable(net.props, net.terms, net.nvars)
    s = L.Store(values=net.store)
    r = t.fixpoint(s)
    assert not r.is_bot and np.array_equal(s.read(), want)
    sd = L.Store(values=net.store)
    rd = t.fixpoint(sd, mode=L.MODE_SWEEP)
    assert np.array_equal(sd.read(), want) and rd.deductions == rd.sweeps * len(net.props)
    assert len(net.props) <= r.deductions < rd.deductions, (r.deductions, rd.deductions)   # tiles at rest were skipped
    r2 = t.fixpoint(s)
    assert not r2.has_changed and r2.sweeps == 1 and r2.deductions == len(net.props)
    buf = net.store.copy()
    t.fixpoint_host(buf)
    assert np.array_equal(buf, want)


def random_pc(rng, nvars):
    from lala_pc_b200 import pcflat
    forms = []
    for _ in range(int(rng.integers(1, 8))):
        kind = int(rng.integers(1, 11))
        vs = rng.permutation(nvars)
        if kind in (1, 2, 7, 8, 9, 10):
            n = int(rng.integers(1, min(6, nvars - 1)))
            ts = [(int(rng.choice([1, 1, 2, 3, 5, -1, -2])), int(v)) for v in vs[:n]]
            if kind == 10 and n == 1 and ts[0][0] == 1:
                ts = [(2, ts[0][1])]
            tree = pcflat.to_tree(kind, ts, int(rng.integers(-10, 40)), int(vs[n]) if kind in (2, 10) else -1)
            if n == 2 and ts[1][0] == -1 and kind != 2 and rng.random() < 0.5:
                # the same constraint written with Unary<Neg> / Binary<Sub> (terms.hpp:87-102, 209-229)
                c0, v0 = ts[0]
                leaf = ("var", v0) if c0 == 1 else ("neg", ("var", v0)) if c0 == -1 else ("mul", ("const", c0), ("var", v0))
                sub = ("sub", leaf, ("var", ts[1][1]))
                tree = tuple(sub if isinstance(e, tuple) and e[0] == "add" else e for e in tree)
            elif n == 1 and ts[0][0] == -1 and kind != 2 and rng.random() < 0.5:
                tree = tuple(("neg", ("var", ts[0][1])) if isinstance(e, tuple) and e[0] == "mul" else e for e in tree)
            forms.append(tree)
        elif kind == 4 and rng.random() < 0.4:
            forms.append(pcflat.to_tree(4, [(1, int(vs[0]))], int(rng.integers(-3, 8)), -1))
        elif kind == 5:
            n = int(rng.integers(1, min(5, nvars)))
            forms.append(pcflat.to_tree(5, [(int(rng.choice([1, -1])), int(v)) for v in vs[:n]], 0, -1))
        else:
            forms.append(pcflat.to_tree(kind, [(1, int(vs[0])), (1, int(vs[1]))], 0, -1))
    return forms


def random_tree_pc(rng, nvars, n_forms=None):
    """Random general formulas - the shapes that have no flat kind and run through the tree interpreter (pc_tree.cuh):
    comparisons between two non-constant terms, min / max / products of variables, nested sums, and / or / equiv /
    imply / xor nests over comparisons and literals."""
    def leaf():
        r = rng.random()
        v = ("var", int(rng.integers(0, nvars)))
        if r < 0.55:
            return v
        if r < 0.70:
            return ("const", int(rng.integers(-4, 9)))
        if r < 0.80:
            return ("neg", v)
        if r < 0.88:
            return ("abs", v)
        return ("mul", ("const", int(rng.choice([2, 3, -1, -2]))), v)

    def term(depth):
        r = rng.random()
        if depth <= 1 or r < 0.35:
            return leaf()
        if r < 0.85:
            op = str(rng.choice(["add", "sub", "mul", "min", "max", "add", "tdiv", "fdiv", "cdiv", "ediv"]))
            return (op, term(depth - 1), term(depth - 1))
        if r < 0.93:
            return (str(rng.choice(["sum", "sum", "prod"])),) + tuple(term(depth - 1) for _ in range(int(rng.integers(2, 5))))
        return (str(rng.choice(["neg", "abs"])), term(depth - 1))

    def formula(depth):
        r = rng.random()
        if depth <= 1 or r < 0.3:
            q = rng.random()
            if q < 0.05:   # the constant formulas (formula.hpp:169-239)
                return ("true",) if rng.random() < 0.6 else ("false",)
            if q < 0.2:
                return (str(rng.choice(["lit", "nlit"])), int(rng.integers(0, nvars)))
            if q < 0.3:
                return ("ae", str(rng.choice(["le", "ge", "eq", "ne"])), int(rng.integers(0, nvars)), int(rng.integers(-3, 9)))
            return (str(rng.choice(["le", "gt", "eq", "ne", "le"])), term(3), term(3))
        op = str(rng.choice(["and", "or", "equiv", "imply", "xor"]))
        return (op, formula(depth - 1), formula(depth - 1))

    return [formula(int(rng.integers(1, 4))) for _ in range(n_forms or int(rng.integers(1, 6)))]


def literal_vars(forms):
    """Variables used as Boolean literals somewhere in the formulas."""
    out = set()

    def walk(f):
        if f[0] in ("lit", "nlit"):
            out.add(int(f[1]))
        elif f[0] in ("and", "or", "equiv", "imply", "xor"):
            walk(f[1])
            walk(f[2])
    for f in forms:
        if isinstance(f, tuple):
            walk(f)
    return sorted(out)


def booleanize(forms, store):
    """A variable used as a literal gets a 0/1 domain. VariableLiteral reads `b does not contain 0` as true but deduces
    b = 1 (formula.hpp:100-120): on a wider domain ask and deduce disagree ([2,2] is "true" yet meets [1,1] to nothing),
    the propagator is not monotone and the fixpoint depends on the evaluation order - an ill-typed model, not a parity
    case (the reference's interpreter only builds literals over Boolean variables)."""
    for v in literal_vars(forms):
        store[v] = (max(int(store[v, 0]), 0), min(int(store[v, 1]), 1)) if store[v, 0] <= 1 and store[v, 1] >= 0 else (0, 1)
    return store


def test_tree_propagators(L, O):
    """Every pc::Formula / pc::Term shape without a flat kind keeps its tree (LPC_PC_TREE) and is walked on the device:
    random formula nests, alone and mixed with flat propagators in one table (tiles + tree list in one fixpoint), plus
    PC::deduce(i) step by step on them."""
    from lala_pc_b200 import pcflat
    rng = np.random.default_rng(2024)
    n_ok = n_tree = 0
    for trial in range(150):
        nvars = int(rng.integers(4, 10))
        forms = random_tree_pc(rng, nvars)
        if trial % 2:
            forms += random_pc(rng, nvars)[:3]
        props, _ = pcflat.flatten(forms)
        n_tree += int((props[:, 0] == 11).sum())
        a = rng.integers(-6, 12, (nvars, 2))
        store = np.stack([a.min(1), a.max(1)], axis=1).astype(np.int32)
        store[rng.random(nvars) < 0.4] = (0, 1)   # finite domains only: x > x + y walks an infinite bound one unit per sweep
        booleanize(forms, store)
        _, r, st = check_parity(L, O, forms, store, f"tree {trial}")
        n_ok += not st.is_bot
        if trial < 40:   # deduce(i) in index order == the tree walker's own steps (return value included)
            m = O.PCModel(forms)
            t = L.PcTable(*pcflat.flatten(forms), len(store))
            s = L.Store(values=store)
            cur, bot = store.copy(), False
            for i in range(len(forms)):
                cur, changed, bot = m.deduce(i, cur, bot)
                assert t.deduce(s, i) == changed, (trial, i, forms[i])
                if not bot:
                    assert np.array_equal(s.read(), cur), (trial, i, forms[i])
    assert n_ok >= 40 and n_tree >= 150
    # the deepest trees the interpreter takes: a left comb of 8 term levels under 6 connective levels
    comb = ("var", 0)
    for k in range(7):
        comb = (("add", "max", "sub")[k % 3], comb, ("var", 1 + k % 3))
    nest = ("le", comb, ("const", 9))
    for k in range(5):
        nest = (("and", "or", "imply", "equiv", "and")[k], nest, ("gt", ("var", k % 4), ("const", k - 2)))
    store = np.array([[0, 5], [-1, 4], [0, 3], [1, 2]], dtype=np.int32)
    check_parity(L, O, [nest, ("le", comb, ("const", 4))], store, "deepest")
    # too deep for the device interpreter: refused, not approximated
    deep = ("var", 0)
    for _ in range(9):
        deep = ("add", deep, ("var", 1))
    with pytest.raises(pcflat.Unsupported):
        pcflat.flatten([("le", deep, ("const", 3))])
    with pytest.raises(L.LpcError):   # a malformed stream never reaches the device
        L.PcTable(np.array([[11, 0, 1, 0, -1]], dtype=np.int32), np.array([[22, 2]], dtype=np.int32), 2)


def test_random_networks(L, O):
    """Small random networks over all flat kinds (negative coefficients, 0/1 and wide domains, a few infinite
    bounds): bot flags and non-failed stores must match the tree walker."""
    rng = np.random.default_rng(33)
    n_ok = 0
    for trial in range(200):
        nvars = int(rng.integers(4, 10))
        forms = random_pc(rng, nvars)
        a = rng.integers(-6, 12, (nvars, 2))
        store = np.stack([a.min(1), a.max(1)], axis=1).astype(np.int32)
        boolish = rng.random(nvars) < 0.4
        store[boolish] = (0, 1)
        if rng.random() < 0.2:
            store[int(rng.integers(0, nvars))] = (-2**31, 2**31 - 1)
        _, r, st = check_parity(L, O, forms, store, f"random {trial}")
        n_ok += not st.is_bot
    assert n_ok >= 40


def test_fallback_paths_of_the_tile_kernel(L, O):
    """Propagators the lane tiles cannot take in int32 go through the term-by-term routine inside the same kernel: sums
    of more than 32 terms, right-hand sides beyond 2^30, term bounds beyond 2^24, infinite and near-infinite domains.
    Mixed with ordinary propagators in one table so that tiles, wild lanes and the long list all run in one fixpoint."""
    from lala_pc_b200 import pcflat
    rng = np.random.default_rng(99)
    NI, PI = -2**31, 2**31 - 1
    n_ok = 0
    for trial in range(60):
        nvars = 48
        a = rng.integers(-4, 9, (nvars, 2))
        store = np.stack([a.min(1), a.max(1)], axis=1).astype(np.int32)
        store[rng.random(nvars) < 0.3] = (0, 1)
        sol = rng.integers(store[:, 0], store[:, 1] + 1)             # a planted assignment keeps most networks alive

        def val(ts):
            return int(sum(c * int(sol[v]) for c, v in ts))
        vs = rng.permutation(nvars)
        forms = []
        # a long sum (40 lanes > 32) and its reified twin
        ts = [(int(rng.choice([1, 1, 2, -1])), int(v)) for v in vs[:40]]
        forms.append(pcflat.to_tree(1, ts, val(ts) + int(rng.integers(0, 6)), -1))
        store[int(vs[41])] = (0, 1)
        forms.append(pcflat.to_tree(2, ts[:36], val(ts[:36]) + int(rng.integers(-3, 4)), int(vs[41])))
        # huge right-hand sides (static fallback) and huge coefficients (dynamic fallback)
        forms.append(pcflat.to_tree(1, [(1, int(vs[0])), (1, int(vs[1]))], int(rng.choice([2**30, 2**30 + 5, PI - 1])), -1))
        big = [(2**23, int(vs[2])), (3, int(vs[3]))]
        forms.append(pcflat.to_tree(7, big, val(big) - int(rng.integers(0, 2**24)), -1))
        eq3 = [(1, int(vs[4])), (-1, int(vs[5])), (1, int(vs[6]))]
        forms.append(pcflat.to_tree(9, eq3, val(eq3), -1))
        forms.append(pcflat.to_tree(10, [(1, int(vs[7])), (2, int(vs[8]))], 0, int(vs[9])))
        store[int(vs[9])] = (NI, PI) if trial % 3 == 0 else (-40, 40)
        forms += random_pc(rng, nvars)[:2]
        for v in vs[42:45]:                                          # infinite, half-infinite and large domains
            store[int(v)] = [(NI, PI), (NI, 7), (-3, PI), (-2**27, 2**27), (2**25, 2**25 + 3)][int(rng.integers(0, 5))]
        _, r, st = check_parity(L, O, forms, store, f"fallback {trial}")
        n_ok += not st.is_bot
    assert n_ok >= 10


def test_errors(L):
    with pytest.raises(L.LpcError):
        L.PcTable(np.array([[99, 0, 1, 0, -1]], dtype=np.int32), np.array([[1, 0]], dtype=np.int32), 2)     # bad kind
    with pytest.raises(L.LpcError):
        L.PcTable(np.array([[1, 0, 2, 0, -1]], dtype=np.int32), np.array([[1, 0]], dtype=np.int32), 2)      # term range
    with pytest.raises(L.LpcError):
        L.PcTable(np.array([[1, 0, 1, 0, -1]], dtype=np.int32), np.array([[1, 5]], dtype=np.int32), 2)      # variable
    with pytest.raises(L.LpcError):
        L.PcTable(np.array([[2, 0, 1, 0, -1]], dtype=np.int32), np.array([[1, 0]], dtype=np.int32), 2)      # no bvar
    t = L.PcTable(np.zeros((0, 5), dtype=np.int32), np.zeros((0, 2), dtype=np.int32), 2)
    s = L.Store(values=np.array([[0, 5], [1, 2]], dtype=np.int32))
    r = t.fixpoint(s)
    assert r.sweeps == 0 and not r.has_changed and not r.is_bot


# ---- bitset stores (VStore<NBitset<64>>, tests/pc_bitset_test.cpp) -------------------------------------------------
def check_parity_bits(L, O, formulas, cells, nvars, label=""):
    from lala_pc_b200 import pcflat
    props, terms = pcflat.flatten(formulas, bitset=True)
    m = O.PCModel(formulas)
    want, st = m.fixpoint_bits(cells)
    t = L.PcTable(props, terms, nvars)
    s = L.Store(nvars=nvars)
    s.write_bits(cells)
    r = t.fixpoint(s, bitset=True)
    got = s.read_bits()
    assert bool(r.is_bot) == bool(st.is_bot), (label, "bot flag")
    if not st.is_bot:
        assert np.array_equal(got, want), (label, "cells", np.flatnonzero(got != want)[:5].tolist())
        n_ent, bits = m.ask_all_bits(want, want_bits=True)
        g_ent, gbits = t.ask_all(s, want_bits=True, bitset=True)
        assert np.array_equal(gbits, bits) and g_ent == n_ent, (label, "ask bits")
        assert bool(r.has_changed) == bool(st.has_changed), (label, "has_changed")
    return got, r, st


def test_bitset_golden_vectors(L, O):
    for k in load_golden("pc_bitset_kat.json")["props"]:
        cells = np.array(k["before_bits"], dtype=np.uint64)
        got, r, st = check_parity_bits(L, O, [to_tree(p) for p in k["props"]], cells, len(cells), k["name"])
        assert got.tolist() == k["after_bits"], k["name"]
        assert bool(r.has_changed) == k["changed"], k["name"]


def test_bitset_deduce_one_and_random(L, O):
    from lala_pc_b200 import pcflat
    from test_devhost_pc import random_bits_pc, random_cells
    rng = np.random.default_rng(5)
    n_ok = 0
    for trial in range(150):
        nvars = int(rng.integers(3, 9))
        forms, cells = random_bits_pc(rng, nvars), random_cells(rng, nvars)
        _, r, st = check_parity_bits(L, O, forms, cells, nvars, f"random bits {trial}")
        n_ok += not st.is_bot
        if trial < 25:   # PC::deduce(i) step by step
            props, terms = pcflat.flatten(forms, bitset=True)
            m = O.PCModel(forms)
            t = L.PcTable(props, terms, nvars)
            s = L.Store(nvars=nvars)
            s.write_bits(cells)
            cur, bot = cells, False
            for i in range(len(forms)):
                cur, changed, bot = m.deduce_bits(i, cur, bot)
                assert t.deduce(s, i, bitset=True) == changed, (trial, i)
                if not bot:
                    assert np.array_equal(s.read_bits(), cur), (trial, i)
    assert n_ok >= 25


@pytest.mark.parametrize("scale", [0.05, 1.0])
def test_config5_interval_vs_bitset(L, O, W, scale):
    """BASELINE.json config 5 (100k vars / 500k propagators at scale 1): both stores bit-exact against the tree walker,
    the bitset fixpoint refines the interval one, the planted solution survives, and the fixpoint is idempotent."""
    net = W.config5(scale)
    formulas = net.formulas()
    m = O.PCModel(formulas)
    t = L.PcTable(net.props, net.terms, net.nvars)
    want_i, st_i = m.fixpoint(net.store)
    si = L.Store(values=net.store)
    ri = t.fixpoint(si)
    assert not ri.is_bot and np.array_equal(si.read(), want_i)
    cells = L.nbit_from_intervals(net.store)
    assert np.array_equal(cells, O.nbit_store(net.store))
    want_b, st_b = m.fixpoint_bits(cells)
    sb = L.Store(nvars=net.nvars)
    sb.write_bits(cells)
    rb = t.fixpoint(sb, bitset=True)
    got_b = sb.read_bits()
    assert not rb.is_bot and not st_b.is_bot and np.array_equal(got_b, want_b)
    sbd = L.Store(nvars=net.nvars)
    sbd.write_bits(cells)
    rbd = t.fixpoint(sbd, bitset=True, mode=L.MODE_SWEEP)
    assert np.array_equal(sbd.read_bits(), want_b) and rb.deductions <= rbd.deductions == rbd.sweeps * len(net.props)
    assert ((got_b & ~L.nbit_from_intervals(want_i)) == 0).all()
    sol = L.nbit_range(net.solution, net.solution)
    assert ((got_b & sol) == sol).all()
    r2 = t.fixpoint(sb, bitset=True)
    assert not r2.has_changed and r2.sweeps == 1
    buf = cells.copy()
    t.fixpoint_host(buf, bitset=True)
    assert np.array_equal(buf, want_b)
    # failing twin: a variable of an equality pushed off its planted value
    eq = np.flatnonzero(net.props[:, 0] == 3)[0]
    v = int(net.terms[net.props[eq, 1], 1])
    twin = cells.copy()
    twin[v] = L.nbit_range((int(net.solution[v]) + 30) % 62, (int(net.solution[v]) + 30) % 62)
    _, stt = m.fixpoint_bits(twin)
    sb.write_bits(twin)
    rt = t.fixpoint(sb, bitset=True)
    assert bool(rt.is_bot) == bool(stt.is_bot)


def test_bitset_linear_and_tree_networks(L, O):
    """NBitset arithmetic: sums of every linear kind, general trees and the four flat shapes in one
    table over a bitset store - the table builder hands the linear kinds to the tree interpreter over the NBitset universe
    (pc_tree.cuh, UNb). Compared with the tree-walking checker's NBitset universe (parity unpinned upstream beyond the
    pc_bitset_test.cpp goldens); fixpoint, ask bits and PC::deduce(i) step by step. The same table still runs on an
    interval store."""
    from lala_pc_b200 import pcflat
    from test_devhost_pc import random_bits_pc, random_cells, small_cells
    rng = np.random.default_rng(29)
    n_ok = n_lin = 0
    for trial in range(160):
        nvars = int(rng.integers(4, 9))
        forms = random_pc(rng, nvars)
        if trial % 2:
            forms += random_tree_pc(rng, nvars)[:2]
        if trial % 5 == 0:
            forms += random_bits_pc(rng, nvars)[:2]
        cells = small_cells(rng, nvars) if trial % 3 else random_cells(rng, nvars)
        for v in literal_vars(forms):
            cells[v] = O.nbit(0, 1)
        props, terms = pcflat.flatten(forms, bitset=True)
        n_lin += int(np.isin(props[:, 0], (1, 2, 7, 8, 9, 10)).sum())
        _, r, st = check_parity_bits(L, O, forms, cells, nvars, f"bits linear/tree {trial}")
        n_ok += not st.is_bot
        if trial < 30:
            m = O.PCModel(forms)
            t = L.PcTable(props, terms, nvars)
            s = L.Store(nvars=nvars)
            s.write_bits(cells)
            cur, bot = cells, False
            for i in range(len(forms)):
                was_bot = bot
                cur, changed, bot = m.deduce_bits(i, cur, bot)
                got = t.deduce(s, i, bitset=True)
                if not was_bot:   # on a failed store the reference's `false` node reports no change (its flag is set already)
                    assert got == changed, (trial, i, forms[i])
                if not bot:
                    assert np.array_equal(s.read_bits(), cur), (trial, i, forms[i])
    assert n_ok >= 30 and n_lin >= 150, (n_ok, n_lin)
    # one table, both universes
    forms = [("le", ("sum", ("var", 0), ("mul", ("const", 2), ("var", 1)), ("var", 2)), ("const", 9)), ("ne", ("var", 0), ("var", 1))]
    props, terms = pcflat.flatten(forms, bitset=True)
    t = L.PcTable(props, terms, 3)
    store = np.array([[0, 9], [3, 3], [2, 8]], dtype=np.int32)
    si = L.Store(values=store)
    t.fixpoint(si)
    want_i, _ = O.PCModel(forms).fixpoint(store)
    assert np.array_equal(si.read(), want_i)
    sb = L.Store(nvars=3)
    sb.write_bits(L.nbit_from_intervals(store))
    t.fixpoint(sb, bitset=True)
    want_b, _ = O.PCModel(forms).fixpoint_bits(L.nbit_from_intervals(store))
    assert np.array_equal(sb.read_bits(), want_b)


def test_bitset_config3_shape(L, O, W):
    """BASELINE.json config 3's propagators (sums and reified sums) over a bitset store, domains inside the universe."""
    net = W.config3(0.02)
    store = np.stack([np.clip(net.store[:, 0], 0, 40), np.clip(net.store[:, 1], 0, 40)], axis=1).astype(np.int32)
    store[store[:, 0] > store[:, 1]] = (0, 40)
    cells = L.nbit_from_intervals(store)
    forms = net.formulas()
    _, r, st = check_parity_bits(L, O, forms, cells, net.nvars, "config3 over bitsets")
    assert r.sweeps >= 1
