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
