"""Pins the CPU oracle (oracle/pir_oracle.cpp) against the reference's own golden vectors and properties.

Sources: tests/pir_test.cpp goldens in record form (tests/golden/pir_kat.json, made by make_pir_kat.py) and the
exhaustive / sampled / singleton bounds-consistency properties of tests/bound_consistency_test.hpp."""
import os

import numpy as np
import pytest

from conftest import load_golden
from oracle import oracle as O

ADD, MUL, MIN, MAX, TDIV, FDIV, CDIV, EDIV, EQ, LEQ = 2, 4, 6, 7, 25, 27, 29, 31, 46, 48
OPS = dict(ADD=ADD, MUL=MUL, MIN=MIN, MAX=MAX, TDIV=TDIV, FDIV=FDIV, CDIV=CDIV, EDIV=EDIV, EQ=EQ, LEQ=LEQ)
COMPLETE = {"EQ", "LEQ", "ADD", "MIN", "MAX"}   # tests/pir_test.cpp:123-132


def run_kat(k, fixpoint):
    store = np.array(k["store"], dtype=np.int32)
    recs = np.array(k["records"], dtype=np.int32)
    if k["clamp"]:
        store = O.pir_clamp_reified(store, recs)
    return fixpoint(store, recs), recs


def test_kat_count(pir_kats):
    assert len(pir_kats) >= 80


def test_golden_vectors(pir_kats):
    for k in pir_kats:
        (s, st), recs = run_kat(k, O.pir_fixpoint)
        if k["bot"]:
            assert st.is_bot and st.has_changed, k["name"]       # deduce_and_test_bot, pir_test.cpp:75-89
            continue
        assert not st.is_bot, k["name"]
        after = np.array(k["after"], dtype=np.int32)
        assert np.array_equal(s[:len(after)], after), (k["name"], s[:len(after)].tolist(), k["after"])
        if k["ua"] is not None:
            ua = O.pir_ask_all(s, recs) == len(recs)             # test_extract, pir_test.cpp:31-52
            assert ua == k["ua"], k["name"]


def test_golden_vectors_any_schedule(pir_kats):
    """The fixpoint must not depend on the order of the propagators (monotone functions)."""
    rng = np.random.default_rng(7)
    for k in pir_kats:
        if k["bot"]:
            continue
        n = len(k["records"])
        for _ in range(3):
            perm = rng.permutation(n)
            (s, st), _ = run_kat(k, lambda a, b: O.pir_fixpoint(a, b, perm=perm))
            after = np.array(k["after"], dtype=np.int32)
            assert np.array_equal(s[:len(after)], after), k["name"]


@pytest.mark.parametrize("name", list(OPS))
def test_exhaustive_bounds_consistency(name):
    """bound_consistency_test.hpp:155-225. Full range [-10,10] with LPC_FULL=1 (46 s on 8 cores; its committed
    result is tests/golden/pir_exhaustive_full.json), [-7,7] otherwise."""
    full = os.environ.get("LPC_FULL") == "1"
    lo, hi = (-10, 10) if full else (-7, 7)
    r, _ = O.pir_exhaustive(OPS[name], lo, hi, name in COMPLETE)
    assert r["cases"] == O.lib().lpco_pir_exhaustive_count(lo, hi)
    for key in ("unsound", "incomplete", "spurious_bot", "bad_entail", "not_converged"):
        assert r[key] == 0, (name, r)
    assert r["bot_cases"] > 0 and r["entailed_cases"] > 0
    if full:
        gold = load_golden("pir_exhaustive_full.json")[name]
        assert {k: r[k] for k in ("cases", "bot_cases", "entailed_cases")} == \
               {k: gold[k] for k in ("cases", "bot_cases", "entailed_cases")}


def test_exhaustive_full_range_record():
    """The committed record of the full [-10,10]^3 run (12,326,391 triples per operator)."""
    gold = load_golden("pir_exhaustive_full.json")
    assert set(gold) == set(OPS)
    for name, r in gold.items():
        assert r["cases"] == 12326391
        assert r["unsound"] == r["incomplete"] == r["spurious_bot"] == r["bad_entail"] == r["not_converged"] == 0


def _np_div(op, a, b):
    """Vectorised battery::{t,f,c,e}div (b != 0), independent of the oracle's C code."""
    bb = np.where(b == 0, 1, b)
    f = a // bb
    r = a - f * bb
    if op == FDIV:
        return f
    if op == CDIV:
        return f + (r != 0)
    if op == TDIV:
        return np.where((r != 0) & ((a < 0) != (bb < 0)), f + 1, f)
    return np.where(bb > 0, f, f + (r != 0))   # EDIV: remainder in [0, |b|)


def _pred(op):
    return {EQ: lambda x, y, z: ((x == 0) | (x == 1)) & (x == (y == z)),
            LEQ: lambda x, y, z: ((x == 0) | (x == 1)) & (x == (y <= z)),
            ADD: lambda x, y, z: x == y + z, MIN: lambda x, y, z: x == np.minimum(y, z),
            MAX: lambda x, y, z: x == np.maximum(y, z), MUL: lambda x, y, z: x == y * z,
            }.get(op, lambda x, y, z: (z != 0) & (x == _np_div(op, y, z)))


@pytest.mark.parametrize("name", list(OPS))
def test_sampled_bounds_consistency(name):
    """bound_consistency_test.hpp:55-113: 23 hand-picked intervals (incl. the empty Itv(1,0)) cubed."""
    op, pred = OPS[name], _pred(OPS[name])
    lo, hi = -20, 20
    itvs = [(lo, lo), (lo + 1, hi - 1), (lo, -3), (lo, -2), (lo, 0), (lo, hi), (-2, -1), (-1, 0), (-2, 2), (-1, -1),
            (0, 0), (1, 1), (2, 2), (0, 1), (1, 2), (0, hi), (2, hi), (3, hi), (hi, hi), (1, 0), (2, 10), (-25, 25),
            (-2, 3)]
    rec = np.array([[op, 0, 1, 2]], dtype=np.int32)
    for x in itvs:
        for y in itvs:
            for z in itvs:
                store = O.pir_clamp_reified(np.array([x, y, z], dtype=np.int32), rec)
                a, b, c = np.meshgrid(np.arange(x[0], x[1] + 1), np.arange(y[0], y[1] + 1),
                                      np.arange(z[0], z[1] + 1), indexing="ij")
                ok = pred(a, b, c)
                has_sol = bool(ok.any())
                s, st = O.pir_fixpoint(store, rec)
                if st.is_bot:
                    assert not has_sol, (name, x, y, z)
                    continue
                if has_sol:
                    hull = np.array([[a[ok].min(), a[ok].max()], [b[ok].min(), b[ok].max()],
                                     [c[ok].min(), c[ok].max()]])
                    if name in COMPLETE:
                        assert np.array_equal(s, hull), (name, x, y, z, s.tolist(), hull.tolist())
                    else:
                        assert (s[:, 0] <= hull[:, 0]).all() and (s[:, 1] >= hull[:, 1]).all(), (name, x, y, z)
                else:
                    assert name not in COMPLETE, (name, x, y, z, s.tolist())
                if O.pir_ask(s, rec[0]):
                    assert has_sol, (name, x, y, z)


@pytest.mark.parametrize("name", list(OPS))
def test_singleton_completeness(name):
    """bound_consistency_test.hpp:115-153: on singletons in [-5,5)^3, ask <=> predicate, and not predicate => bot."""
    op, pred = OPS[name], _pred(OPS[name])
    rec = np.array([[op, 0, 1, 2]], dtype=np.int32)
    for i in range(-5, 5):
        for j in range(-5, 5):
            for k in range(-5, 5):
                store = O.pir_clamp_reified(np.array([[i, i], [j, j], [k, k]], dtype=np.int32), rec)
                ent = O.pir_ask(store, rec[0])
                s, st = O.pir_fixpoint(store, rec)
                if bool(pred(np.int64(i), np.int64(j), np.int64(k))):
                    assert ent and not st.is_bot and np.array_equal(s, store), (name, i, j, k)
                else:
                    assert not ent and st.is_bot, (name, i, j, k)


def test_division_helpers():
    """battery::{t,f,c,e}div as used for the concrete semantics at pir_test.cpp:129-132."""
    for a in range(-25, 26):
        for b in list(range(-7, 0)) + list(range(1, 8)):
            assert O.div(a, TDIV, b) == int(a / b)
            assert O.div(a, FDIV, b) == a // b
            assert O.div(a, CDIV, b) == -((-a) // b)
            q = O.div(a, EDIV, b)
            r = a - q * b
            assert 0 <= r < abs(b)


def test_ternary_div_hull():
    """PIRTest.TernaryDiv (pir_test.cpp:91-118): x = y ediv z on x,y in [-4,4], z = -5; soundness only."""
    rec = np.array([[EDIV, 0, 1, 2]], dtype=np.int32)
    store = np.array([[-4, 4], [-4, 4], [-5, -5]], dtype=np.int32)
    sols = [(a, b, -5) for a in range(-4, 5) for b in range(-4, 5) if a == O.div(b, EDIV, -5)]
    h = np.array(sols)
    s, st = O.pir_fixpoint(store, rec)
    assert not st.is_bot
    assert (s[:, 0] <= h.min(0)).all() and (s[:, 1] >= h.max(0)).all()


def test_failed_store_stops_at_bot():
    """TopProp (pir_test.cpp:322-329): bounds of a failed store keep moving; the contract is stop at first bot."""
    k = [c for c in load_golden("pir_kat.json") if c["name"] == "TopProp"][0]
    s, st = O.pir_fixpoint(np.array(k["store"], dtype=np.int32), np.array(k["records"], dtype=np.int32))
    assert st.is_bot and st.sweeps <= 3
