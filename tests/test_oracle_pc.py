"""Pins the PC part of the CPU oracle (oracle/pc_oracle.cpp) against the reference's golden vectors
(tests/pc_test.cpp; record form in tests/golden/pc_kat.json, made by make_pc_kat.py)."""
import numpy as np

from conftest import load_golden
from oracle import oracle as O


def to_tree(x):
    return tuple(to_tree(e) if isinstance(e, list) else e for e in x)


def test_term_goldens():
    """TermTest.AddTermBinary / AddTermNary (pc_test.cpp:31-67): project, embed, project again."""
    for k in load_golden("pc_kat.json")["terms"]:
        store = np.array(k["store"], dtype=np.int32)
        term = to_tree(k["term"])
        assert list(O.pc_term_project(term, store)) == k["project"], k["name"]
        store2, changed = O.pc_term_embed(term, store, *k["embed"])
        assert changed == k["changed"], k["name"]
        assert list(O.pc_term_project(term, store2)) == k["project_after"], k["name"]


def test_formula_goldens():
    kats = load_golden("pc_kat.json")["props"]
    assert len(kats) >= 50
    for k in kats:
        m = O.PCModel([to_tree(p) for p in k["props"]])
        store = np.array(k["store"], dtype=np.int32)
        # the reference's tests iterate without a stop condition (pc_test.cpp:85-100)
        s, st = m.fixpoint(store, stop_on_bot=False, max_sweeps=1000)
        assert st.sweeps < 1000, k["name"]
        if k["bot"]:
            assert st.is_bot, k["name"]
            assert all(lb > ub for lb, ub in s[:len(k["store"])]), (k["name"], s.tolist())
            continue
        assert not st.is_bot, k["name"]
        after = np.array(k["after"], dtype=np.int32)
        assert np.array_equal(s[:len(after)], after), (k["name"], s.tolist(), k["after"])
        if k["changed"] is not None:
            assert bool(st.has_changed) == k["changed"], k["name"]
        if k["ua"] is not None:
            assert (m.ask_all(s) == len(m)) == k["ua"], k["name"]
        # stop-on-bot contract gives the same store on non-failed inputs
        s2, st2 = m.fixpoint(store)
        assert np.array_equal(s2, s) and not st2.is_bot


def test_bitset_goldens():
    """BitPCTest.* (pc_bitset_test.cpp:68-157): the tree walker over NBitset<64> cells."""
    kats = load_golden("pc_bitset_kat.json")["props"]
    assert len(kats) == 15
    for k in kats:
        m = O.PCModel([to_tree(p) for p in k["props"]])
        before = np.array(k["before_bits"], dtype=np.uint64)
        s, st = m.fixpoint_bits(before, stop_on_bot=False, max_sweeps=100)
        assert not st.is_bot and st.sweeps < 100, k["name"]
        assert s.tolist() == k["after_bits"], (k["name"], [hex(int(x)) for x in s])
        assert bool(st.has_changed) == k["changed"], k["name"]
        assert (m.ask_all_bits(s) == len(m)) == k["ua"], k["name"]


def test_nbit_layout():
    """NBit(lb, ub), from_set, lb()/ub(): the layout pinned by `var -15..5` == NBit(-1, 5) (pc_bitset_test.cpp:152-156)."""
    assert O.nbit(-15, 5) == O.nbit(-1, 5) == 0b1111111
    assert O.nbit(0, 61) == ((1 << 63) - 2) and O.nbit(-2**31, 2**31 - 1) == 2**64 - 1
    assert O.nbit(62, 100) == 1 << 63 and O.nbit(5, 4) == 0
    assert O.nbit_from_set([1, 3]) == 0b10100
    assert O.nbit_bounds(O.nbit(3, 9)) == (3, 9)
    assert O.nbit_bounds(O.nbit(-5, 70)) == (-2**31, 2**31 - 1)
    assert O.nbit_bounds(1 << 63) == (62, 2**31 - 1) and O.nbit_bounds(1) == (-2**31, -1)
    import lala_pc_b200 as L     # the product's host helper builds the same cells
    lo = np.array([-15, 0, 62, 5, 3, -2**31])
    hi = np.array([5, 61, 100, 4, 9, 2**31 - 1])
    assert L.nbit_range(lo, hi).tolist() == [O.nbit(a, b) for a, b in zip(lo.tolist(), hi.tolist())]


def test_bitset_refines_interval():
    """On domains inside [0, 61] the bitset fixpoint is at least as tight as the interval one (config-5 shapes)."""
    from lala_pc_b200 import workloads as W
    net = W.config5(0.02)
    m = O.PCModel(net.formulas())
    si, sti = m.fixpoint(net.store)
    sb, stb = m.fixpoint_bits(O.nbit_store(net.store))
    assert not sti.is_bot and not stb.is_bot
    hull = O.nbit_store(si)
    assert ((sb & ~hull) == 0).all()
    assert (sb != hull).any()        # != punches holes an interval cannot hold
    sol = np.array([O.nbit(v, v) for v in net.solution.tolist()], dtype=np.uint64)
    assert ((sb & sol) == sol).all()  # the planted solution survives


def test_true_false_nodes():
    """formula.hpp:169-239: True is entailed and does nothing, False fails the element; under connectives they act as
    the constants they are (b <=> true fixes b, false \\/ c enforces c, b => false refutes b)."""
    from oracle import oracle as O
    store = np.array([[0, 1], [0, 10], [0, 1]], dtype=np.int32)
    forms = [("equiv", ("lit", 0), ("true",)), ("or", ("false",), ("le", ("var", 1), ("const", 4))),
             ("imply", ("lit", 2), ("false",)), ("true",)]
    m = O.PCModel(forms)
    out, st = m.fixpoint(store)
    assert not st.is_bot and out.tolist() == [[1, 1], [0, 4], [0, 0]]
    n, bits = m.ask_all(out, want_bits=True)
    assert bits.tolist() == [1, 1, 1, 1]
    _, st = O.PCModel([("false",)]).fixpoint(store)
    assert st.is_bot
    _, st = O.PCModel([("and", ("true",), ("equiv", ("lit", 0), ("false",))), ("lit", 0)]).fixpoint(store)
    assert st.is_bot
