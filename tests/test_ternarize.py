"""The ternarisation front-end (lala-pc_b200/ternarize.py): the ternary PIR network it emits has
exactly the solutions of the formula on the formula's variables. Checked by brute force against a direct evaluation of
the formula, through the CPU restatement's search (not gpu) and through the device search (gpu)."""
import itertools

import numpy as np
import pytest

from oracle import oracle as O


def evaluate(t, a):
    """Direct evaluation of a term / formula under the assignment `a` (list of ints)."""
    op = t[0]
    if op == "var": return a[t[1]]
    if op == "const": return t[1]
    if op == "lit": return int(a[t[1]] != 0)
    if op == "nlit": return int(a[t[1]] == 0)
    if op == "true": return 1
    if op == "false": return 0
    if op == "neg": return -evaluate(t[1], a)
    if op == "abs": return abs(evaluate(t[1], a))
    if op == "not": return 1 - evaluate(t[1], a)
    if op == "sum": return sum(evaluate(s, a) for s in t[1:])
    x, y = evaluate(t[1], a), evaluate(t[2], a)
    return {"add": lambda: x + y, "sub": lambda: x - y, "mul": lambda: x * y, "min": lambda: min(x, y), "max": lambda: max(x, y),
            "le": lambda: int(x <= y), "lt": lambda: int(x < y), "ge": lambda: int(x >= y), "gt": lambda: int(x > y),
            "eq": lambda: int(x == y), "ne": lambda: int(x != y), "and": lambda: int(bool(x) and bool(y)),
            "or": lambda: int(bool(x) or bool(y)), "imply": lambda: int((not x) or bool(y)), "equiv": lambda: int(bool(x) == bool(y))}[op]()


def random_term(rng, nvars, depth):
    if depth == 0 or rng.random() < 0.3:
        return ("var", int(rng.integers(0, nvars))) if rng.random() < 0.75 else ("const", int(rng.integers(-3, 4)))
    op = str(rng.choice(["add", "sub", "mul", "min", "max", "neg", "abs", "sum"]))
    if op in ("neg", "abs"):
        return (op, random_term(rng, nvars, depth - 1))
    if op == "sum":
        return ("sum",) + tuple(random_term(rng, nvars, depth - 1) for _ in range(3))
    return (op, random_term(rng, nvars, depth - 1), random_term(rng, nvars, depth - 1))


def random_formula(rng, nvars, depth):
    if depth == 0 or rng.random() < 0.5:
        return (str(rng.choice(["le", "lt", "ge", "gt", "eq", "ne"])), random_term(rng, nvars, 2), random_term(rng, nvars, 1))
    op = str(rng.choice(["and", "or", "imply", "equiv", "not"]))
    if op == "not":
        return ("not", random_formula(rng, nvars, depth - 1))
    return (op, random_formula(rng, nvars, depth - 1), random_formula(rng, nvars, depth - 1))


def model(rng):
    nvars = int(rng.integers(2, 5))
    a = rng.integers(-3, 4, (nvars, 2))
    store = np.stack([a.min(1), a.max(1)], axis=1).astype(np.int32)
    formulas = [random_formula(rng, nvars, 2) for _ in range(int(rng.integers(1, 4)))]
    want = sum(all(evaluate(f, list(asg)) for f in formulas)
               for asg in itertools.product(*[range(lo, hi + 1) for lo, hi in store.tolist()]))
    return nvars, store, formulas, want


def padded(recs, ext):
    if len(ext) % 2:     # batched stores need an even number of variables
        ext = np.concatenate([ext, np.array([[0, 0]], dtype=np.int32)])
    return recs, ext


def test_known_decompositions():
    from lala_pc_b200 import ternarize as T
    # x + y <= 5 asserted: one temporary, one LEQ with the constant ONE as its result
    recs, ext, n = T.ternarize([("le", ("add", ("var", 0), ("var", 1)), ("const", 5))], [[0, 10], [0, 10]])
    assert n == 2 and len(recs) == 2
    assert recs[0].tolist()[0] == T.ADD and recs[1].tolist()[0] == T.LEQ
    one, five = recs[1][1], recs[1][3]
    assert ext[one].tolist() == [1, 1] and ext[five].tolist() == [5, 5]
    s, st = O.pir_fixpoint(ext, recs)
    assert s[:2].tolist() == [[0, 5], [0, 5]]                    # the result of pir_test.cpp's TemporalConstraint1
    # z = x * y is already ternary: no temporary, no reification
    recs, ext, n = T.ternarize([("eq", ("var", 2), ("mul", ("var", 0), ("var", 1)))], [[0, 3]] * 3)
    assert recs.tolist() == [[T.MUL, 2, 0, 1]] and len(ext) == 3
    # identical sub-terms share their temporary
    recs, ext, n = T.ternarize([("le", ("add", ("var", 0), ("var", 1)), ("const", 5)),
                                ("ge", ("add", ("var", 0), ("var", 1)), ("const", 2))], [[0, 10], [0, 10]])
    assert sum(r[0] == T.ADD for r in recs.tolist()) == 1


def test_solutions_preserved_cpu():
    from lala_pc_b200 import ternarize as T
    rng = np.random.default_rng(2024)
    total = 0
    for trial in range(120):
        nvars, store, formulas, want = model(rng)
        recs, ext, n = T.ternarize(formulas, store)
        recs, ext = padded(recs, ext)
        out = O.pir_search(ext[None], recs, list(range(n)))[0]
        assert out[0] == want, (trial, formulas, store.tolist(), out.tolist(), want)
        assert out[4] == 0 and out[5] == 0, (trial, "every leaf must be decided", out.tolist())
        total += want
    assert total > 500


@pytest.mark.gpu
def test_solutions_preserved_device_search():
    import lala_pc_b200 as L
    from lala_pc_b200 import ternarize as T
    L.device_init(0)
    rng = np.random.default_rng(7)
    for trial in range(40):
        nvars, store, formulas, want = model(rng)
        recs, ext, n = T.ternarize(formulas, store)
        recs, ext = padded(recs, ext)
        t = L.Table(recs, len(ext))
        b = L.Batch(t, 1)
        b.write(ext[None])
        r, per = b.search(list(range(n)))
        assert r.n_solutions == want and r.n_incomplete == 0 and r.n_unknown_leaves == 0, (trial, formulas)
        assert np.array_equal(per, O.pir_search(ext[None], recs, list(range(n))))
        b.close()
