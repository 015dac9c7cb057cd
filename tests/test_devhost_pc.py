"""The flat PC device rules (lala-pc_b200/csrc/pc_device.cuh) compiled for the HOST and compared with the tree-walking
oracle: goldens, the config-3 / config-5 shapes and random networks. A CPU-side check of the code the PC kernel runs."""
import ctypes

import numpy as np
import pytest

from conftest import load_golden
from oracle import oracle as O
from test_devhost import devhost  # noqa: F401  (fixture: builds tests/native/libdevhost.so)
from test_gpu_pc import booleanize, literal_vars, random_pc, random_tree_pc, to_tree


def host_fixpoint(D, props, terms, store):
    s = np.ascontiguousarray(store, dtype=np.int32).copy()
    p = np.ascontiguousarray(props, dtype=np.int32)
    t = np.ascontiguousarray(terms if len(terms) else np.zeros((1, 2)), dtype=np.int32)
    bot, ch = ctypes.c_int(0), ctypes.c_int(0)
    D.devhost_pc_fixpoint(p.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(len(p)), t.ctypes.data_as(ctypes.c_void_p),
                          s.ctypes.data_as(ctypes.c_void_p), len(s), ctypes.byref(bot), ctypes.byref(ch))
    return s, bool(bot.value), bool(ch.value)


def host_ask_bits(D, props, terms, store):
    p = np.ascontiguousarray(props, dtype=np.int32)
    t = np.ascontiguousarray(terms, dtype=np.int32)
    s = np.ascontiguousarray(store, dtype=np.int32)
    return np.array([D.devhost_pc_ask(p.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(i), t.ctypes.data_as(ctypes.c_void_p),
                                      s.ctypes.data_as(ctypes.c_void_p)) for i in range(len(p))], dtype=np.uint8)


def compare(D, formulas, store, label):
    from lala_pc_b200 import pcflat
    props, terms = pcflat.flatten(formulas)
    m = O.PCModel(formulas)
    want, st = m.fixpoint(store)
    got, bot, ch = host_fixpoint(D, props, terms, store)
    assert bot == bool(st.is_bot), label
    if not bot:
        assert np.array_equal(got, want), (label, got.tolist(), want.tolist())
        assert ch == bool(st.has_changed), label
        _, bits = m.ask_all(want, want_bits=True)
        assert np.array_equal(host_ask_bits(D, props, terms, want), bits), label
    return st


def test_goldens_and_flattener(devhost):
    from lala_pc_b200 import pcflat
    ran = n_tree = 0
    for k in load_golden("pc_kat.json")["props"]:
        formulas = [to_tree(p) for p in k["props"]]
        props, _ = pcflat.flatten(formulas)      # flat kinds where there is one, LPC_PC_TREE otherwise
        n_tree += int((props[:, 0] == pcflat.PC_TREE).any())
        ran += 1
        compare(devhost, formulas, np.array(k["store"], dtype=np.int32), k["name"])
    assert ran >= 96 and n_tree >= 53
    # shapes without a flat kind are never approximated by one: strict flattening refuses them
    for f in (("le", ("add", ("add", ("var", 0), ("var", 1)), ("var", 2)), ("const", 3)),
              ("le", ("var", 0), ("var", 1)), ("gt", ("var", 0), ("var", 1)), ("le", ("var", 0), ("add", ("const", -5), ("var", 1))),
              ("le", ("sub", ("var", 0), ("mul", ("const", 3), ("var", 1))), ("const", 2)),
              ("equiv", ("lit", 0), ("and", ("lit", 1), ("lit", 2)))):
        with pytest.raises(pcflat.Unsupported):
            pcflat.flatten([f], tree=False)
        assert pcflat.flatten([f])[0][0, 0] == pcflat.PC_TREE


def test_random_tree_networks(devhost):
    """The tree interpreter (pc_tree.cuh) against the tree-walking checker on random formula nests."""
    rng = np.random.default_rng(17)
    n_ok = 0
    for trial in range(400):
        nvars = int(rng.integers(3, 9))
        forms = random_tree_pc(rng, nvars)
        if trial % 3 == 0:
            forms += random_pc(rng, max(nvars, 4))[:2] if nvars >= 4 else []
        a = rng.integers(-6, 12, (nvars, 2))
        store = np.stack([a.min(1), a.max(1)], axis=1).astype(np.int32)
        store[rng.random(nvars) < 0.4] = (0, 1)   # finite domains only: x > x + y walks an infinite bound one unit per sweep
        booleanize(forms, store)
        st = compare(devhost, forms, store, f"tree {trial}")
        n_ok += not st.is_bot
    assert n_ok >= 100
    # the deepest trees the interpreter takes (8 term levels under 6 connective levels), and one level more: refused
    from lala_pc_b200 import pcflat
    comb = ("var", 0)
    for k in range(7):
        comb = (("add", "max", "sub")[k % 3], comb, ("var", 1 + k % 3))
    nest = ("le", comb, ("const", 9))
    for k in range(5):
        nest = (("and", "or", "imply", "equiv", "and")[k], nest, ("gt", ("var", k % 4), ("const", k - 2)))
    store = np.array([[0, 5], [-1, 4], [0, 3], [1, 2]], dtype=np.int32)
    compare(devhost, [nest, ("le", comb, ("const", 4))], store, "deepest")
    with pytest.raises(pcflat.Unsupported):
        pcflat.flatten([("le", ("add", comb, ("var", 1)), ("const", 4))])
    with pytest.raises(pcflat.Unsupported):
        pcflat.flatten([("and", nest, ("lit", 3))])


@pytest.mark.parametrize("cfg", ["config3", "config5"])
def test_config_shapes(devhost, cfg):
    from lala_pc_b200 import workloads as W
    net = getattr(W, cfg)(0.05)
    st = compare(devhost, net.formulas(), net.store, cfg)
    assert not st.is_bot and st.sweeps >= 3


def test_random_networks(devhost):
    rng = np.random.default_rng(7)
    n_ok = 0
    for trial in range(300):
        nvars = int(rng.integers(4, 10))
        forms = random_pc(rng, nvars)
        a = rng.integers(-6, 12, (nvars, 2))
        store = np.stack([a.min(1), a.max(1)], axis=1).astype(np.int32)
        store[rng.random(nvars) < 0.4] = (0, 1)
        if rng.random() < 0.2:
            store[int(rng.integers(0, nvars))] = (-2**31, 2**31 - 1)
        st = compare(devhost, forms, store, f"random {trial}")
        n_ok += not st.is_bot
    assert n_ok >= 60


# ---- the bitset rules (NBitset<64> cells) ---------------------------------------------------------------------------
def host_fixpoint_bits(D, props, terms, cells):
    s = np.ascontiguousarray(cells, dtype=np.uint64).copy()
    p = np.ascontiguousarray(props if len(props) else np.zeros((1, 5)), dtype=np.int32)
    t = np.ascontiguousarray(terms if len(terms) else np.zeros((1, 2)), dtype=np.int32)
    bot, ch = ctypes.c_int(0), ctypes.c_int(0)
    D.devhost_pc_fixpoint_bits(p.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(len(props)), t.ctypes.data_as(ctypes.c_void_p),
                               ctypes.c_longlong(len(terms)), s.ctypes.data_as(ctypes.c_void_p), len(s), ctypes.byref(bot),
                               ctypes.byref(ch))
    return s, bool(bot.value), bool(ch.value)


def compare_bits(D, formulas, cells, label):
    from lala_pc_b200 import pcflat
    props, terms = pcflat.flatten(formulas, bitset=True)
    m = O.PCModel(formulas)
    want, st = m.fixpoint_bits(cells)
    got, bot, ch = host_fixpoint_bits(D, props, terms, cells)
    assert bot == bool(st.is_bot), label
    if not bot:
        assert np.array_equal(got, want), (label, got.tolist(), want.tolist())
        assert ch == bool(st.has_changed), label
        _, bits = m.ask_all_bits(want, want_bits=True)
        p = np.ascontiguousarray(props, dtype=np.int32)
        t = np.ascontiguousarray(terms if len(terms) else np.zeros((1, 2)), dtype=np.int32)
        w = np.ascontiguousarray(want)
        gb = [D.devhost_pc_ask_bits(p.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(len(p)), ctypes.c_longlong(i),
                                    t.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(len(terms)),
                                    w.ctypes.data_as(ctypes.c_void_p)) for i in range(len(p))]
        assert gb == bits.tolist(), label
    return got, st


def random_bits_pc(rng, nvars):
    from lala_pc_b200 import pcflat
    forms = []
    for _ in range(int(rng.integers(1, 8))):
        kind = int(rng.integers(3, 7))
        vs = rng.permutation(nvars)
        if kind == 4 and rng.random() < 0.4:
            forms.append(pcflat.to_tree(4, [(1, int(vs[0]))], int(rng.integers(-3, 70)), -1))
        elif kind == 5:
            n = int(rng.integers(1, min(5, nvars)))
            forms.append(pcflat.to_tree(5, [(int(rng.choice([1, -1])), int(v)) for v in vs[:n]], 0, -1))
        else:
            forms.append(pcflat.to_tree(kind, [(1, int(vs[0])), (1, int(vs[1]))], 0, -1))
    return forms


def random_cells(rng, nvars):
    cells = np.zeros(nvars, dtype=np.uint64)
    for i in range(nvars):
        r = rng.random()
        if r < 0.35:
            cells[i] = O.nbit(0, 1) if rng.random() < 0.7 else O.nbit(int(rng.integers(0, 2)), 1)
        elif r < 0.7:
            a, b = sorted(int(x) for x in rng.integers(-3, 66, 2))
            cells[i] = O.nbit(a, b)
        else:
            cells[i] = O.nbit_from_set(int(x) for x in rng.integers(-2, 64, int(rng.integers(1, 5))))
    return cells


def test_bitset_goldens(devhost):
    for k in load_golden("pc_bitset_kat.json")["props"]:
        got, st = compare_bits(devhost, [to_tree(p) for p in k["props"]], np.array(k["before_bits"], dtype=np.uint64), k["name"])
        assert got.tolist() == k["after_bits"], k["name"]


def test_bitset_config5_and_random(devhost):
    from lala_pc_b200 import workloads as W
    net = W.config5(0.05)
    _, st = compare_bits(devhost, net.formulas(), O.nbit_store(net.store), "config5 bits")
    assert not st.is_bot and st.sweeps >= 3
    rng = np.random.default_rng(11)
    n_ok = 0
    for trial in range(400):
        nvars = int(rng.integers(3, 9))
        _, st = compare_bits(devhost, random_bits_pc(rng, nvars), random_cells(rng, nvars), f"random bits {trial}")
        n_ok += not st.is_bot
    assert n_ok >= 100


# ---- NBitset arithmetic: sums and general trees over bitset cells (PARITY UNPINNED upstream, see oracle/pc_oracle.cpp) -----
def test_linear_props_as_tree_streams(devhost):
    """The table builder's rewriting of a flat linear propagator (pc_linear_tree_words) == the stream of the formula tree
    the propagator stands for (pcflat.to_tree -> encode_tree)."""
    from lala_pc_b200 import pcflat
    rng = np.random.default_rng(3)
    for _ in range(300):
        kind = int(rng.choice([1, 2, 7, 8, 9, 10]))
        n = int(rng.integers(1, 7))
        ts = [(int(rng.choice([1, 1, 2, 3, -1, -4])), int(rng.integers(0, 9))) for _ in range(n)]
        if kind == 10 and n == 1 and ts[0][0] == 1:
            ts = [(2, ts[0][1])]
        rhs, bvar = int(rng.integers(-20, 60)), int(rng.integers(0, 9))
        _, pairs, _, _ = pcflat.encode_tree(pcflat.to_tree(kind, ts, rhs, bvar))
        want = [w for pr in pairs for w in pr]
        t = np.array(ts, dtype=np.int32)
        out = np.zeros(64, dtype=np.int32)
        k = devhost.devhost_linear_tree_words(kind, t.ctypes.data_as(ctypes.c_void_p), n, rhs, bvar, out.ctypes.data_as(ctypes.c_void_p), 64)
        assert out[:k].tolist() == want, (kind, ts, rhs, bvar)


def small_cells(rng, nvars):
    """Domains inside the exact range of the universe, some with holes."""
    cells = np.zeros(nvars, dtype=np.uint64)
    for i in range(nvars):
        r = rng.random()
        if r < 0.3:
            cells[i] = O.nbit(0, 1)
        elif r < 0.75:
            a, b = sorted(int(x) for x in rng.integers(-1, 20, 2))
            cells[i] = O.nbit(a, b)
        else:
            cells[i] = O.nbit_from_set(int(x) for x in rng.integers(0, 24, int(rng.integers(1, 5))))
    return cells


def test_bitset_linear_and_tree_networks(devhost):
    """Sums (every linear kind), general trees and the four flat shapes mixed, over NBitset cells: the device rules on the
    host against the tree-walking checker's NBitset universe."""
    rng = np.random.default_rng(23)
    n_ok = n_lin = 0
    for trial in range(500):
        nvars = int(rng.integers(4, 9))
        forms = random_pc(rng, nvars)
        if trial % 2:
            forms += random_tree_pc(rng, nvars)[:2]
        if trial % 5 == 0:
            forms += random_bits_pc(rng, nvars)[:2]
        cells = small_cells(rng, nvars) if trial % 3 else random_cells(rng, nvars)
        for v in literal_vars(forms):
            cells[v] = O.nbit(0, 1)
        n_lin += sum(1 for f in forms if f[0] in ("le", "gt", "eq", "equiv"))
        _, st = compare_bits(devhost, forms, cells, f"bits linear/tree {trial}")
        n_ok += not st.is_bot
    assert n_ok >= 100 and n_lin >= 500, (n_ok, n_lin)
