"""Writes tests/golden/pir_kat.json: the reference's PIR golden vectors in record form.

The expected intervals are the ones asserted by the reference's own tests (tests/pir_test.cpp, line numbers in
`source`). The reference builds its propagators from FlatZinc through an un-vendored parser + ternariser, so
each case is hand-ternarised here: `int_plus(a,b,c)` = record (ADD, X=c, Y=a, Z=b), `int_times`, `int_min`,
`int_max`, `int_Xdiv` alike (bound_consistency_test.hpp:67-68 writes `constraint pred(y, z, x)`); constants are
variables with singleton domains; `int_eq(a,b)` = `ONE = (a == b)`, `int_ne` = `ZERO = (a == b)`,
`int_gt(a,b)` = `ZERO = (a <= b)`, `int_lt(a,b)` = `ZERO = (b <= a)`, `a - b = t` = `a = t + b`,
`-a = t` = `ZERO = t + a`, `and` = MIN, `or` = MAX over 0/1 variables. Only the user variables (the first
len(after) store entries) are compared; auxiliary variables follow them in the store.

Run:  python tests/golden/make_pir_kat.py
"""
import json
import os

ADD, MUL, MIN, MAX, TDIV, FDIV, CDIV, EDIV, EQ, LEQ = 2, 4, 6, 7, 25, 27, 29, 31, 46, 48
NI, PI = -2**31, 2**31 - 1
TOP = [NI, PI]

K = []


def kat(name, source, store, records, after=None, bot=False, ua=None, clamp=True):
    """store: initial intervals; after: expected prefix of the store after the fixpoint (None if bot)."""
    K.append(dict(name=name, source=source, store=store, records=records, after=after, bot=bot, ua=ua,
                  clamp=clamp))


# ---- single-record cases ---------------------------------------------------------------------------------------
kat("TernaryProblem", "pir_test.cpp:175-182", [[0, 10], [0, 10], [5, 5]], [[ADD, 2, 0, 1]],
    [[0, 5], [0, 5], [5, 5]], ua=False)
kat("AddEquality", "pir_test.cpp:208-214", [[0, 10], [0, 10], [5, 5]], [[ADD, 2, 0, 1]], [[0, 5], [0, 5]], ua=False)
kat("TemporalConstraint1Flat", "pir_test.cpp:217-224", [[0, 10], [0, 10], [NI, 5]], [[ADD, 2, 0, 1]],
    [[0, 5], [0, 5], [0, 5]], ua=False)
kat("ReifiedEquality", "pir_test.cpp:184-188", [[0, 1], TOP, [2, 2]], [[EQ, 0, 1, 2]], [[0, 1], TOP], ua=False)
kat("MinConstraint1.a", "pir_test.cpp:626-630", [[0, 4], [2, 5], [0, 10]], [[MIN, 2, 0, 1]],
    [[0, 4], [2, 5], [0, 4]], ua=False)
kat("MinConstraint1.b", "pir_test.cpp:632-633", [[0, 4], [2, 5], [0, 3]], [[MIN, 2, 0, 1]],
    [[0, 4], [2, 5], [0, 3]], ua=False)
kat("MinConstraint1.c", "pir_test.cpp:635-636", [[0, 1], [2, 5], [0, 3]], [[MIN, 2, 0, 1]],
    [[0, 1], [2, 5], [0, 1]], ua=False)
kat("MinConstraint1.d", "pir_test.cpp:638-639", [[0, 0], [2, 5], [0, 1]], [[MIN, 2, 0, 1]],
    [[0, 0], [2, 5], [0, 0]], ua=True)
kat("MinConstraint2.c", "pir_test.cpp:652-653", [[4, 4], [2, 5], [0, 3]], [[MIN, 2, 0, 1]],
    [[4, 4], [2, 3], [2, 3]], ua=False)
kat("MaxConstraint1.a", "pir_test.cpp:671-675", [[0, 4], [2, 5], [0, 10]], [[MAX, 2, 0, 1]],
    [[0, 4], [2, 5], [2, 5]], ua=False)
kat("MaxConstraint1.b", "pir_test.cpp:677-678", [[0, 4], [2, 5], [2, 3]], [[MAX, 2, 0, 1]],
    [[0, 3], [2, 3], [2, 3]], ua=False)
kat("MaxConstraint1.c", "pir_test.cpp:680-681", [[0, 1], [2, 3], [2, 3]], [[MAX, 2, 0, 1]],
    [[0, 1], [2, 3], [2, 3]], ua=False)
kat("MaxConstraint1.d", "pir_test.cpp:683-684", [[0, 1], [2, 2], [2, 3]], [[MAX, 2, 0, 1]],
    [[0, 1], [2, 2], [2, 2]], ua=True)
kat("MaxConstraint2.b", "pir_test.cpp:694-695", [[0, 4], [2, 5], [5, 5]], [[MAX, 2, 0, 1]],
    [[0, 4], [5, 5], [5, 5]], ua=True)
for nm, src, before, after, ua in [
    ("IntTimes1.a", "pir_test.cpp:759-766", [[0, 1], [0, 1], [0, 1]], [[0, 1], [0, 1], [0, 1]], False),
    ("IntTimes1.b", "pir_test.cpp:767-768", [[1, 1], [0, 1], [0, 1]], [[1, 1], [0, 1], [0, 1]], False),
    ("IntTimes1.c", "pir_test.cpp:769-770", [[1, 1], [1, 1], [0, 1]], [[1, 1], [1, 1], [1, 1]], True),
    ("IntTimes2", "pir_test.cpp:773-783", [[0, 1], [0, 1], [1, 1]], [[1, 1], [1, 1], [1, 1]], True),
    ("IntTimes3", "pir_test.cpp:785-793", [[0, 0], [0, 1], [0, 1]], [[0, 0], [0, 1], [0, 0]], True),
    ("IntTimes4", "pir_test.cpp:795-803", [[0, 1], [0, 0], [0, 1]], [[0, 1], [0, 0], [0, 0]], True),
    ("IntTimes5", "pir_test.cpp:805-813", [[1, 2], [0, 1], [0, 0]], [[1, 2], [0, 0], [0, 0]], True),
    ("IntTimes6", "pir_test.cpp:815-823", [[0, 1], [1, 2], [0, 0]], [[0, 0], [1, 2], [0, 0]], True),
]:
    kat(nm, src, before, [[MUL, 2, 0, 1]], after, ua=ua)
kat("IntDiv1.a", "pir_test.cpp:825-832", [[0, 1], [0, 1], [0, 1]], [[TDIV, 2, 0, 1]], [[0, 1], [1, 1], [0, 1]],
    ua=False)
kat("IntDiv1.b", "pir_test.cpp:833-834", [[1, 1], [1, 1], [0, 1]], [[TDIV, 2, 0, 1]], [[1, 1], [1, 1], [1, 1]],
    ua=True)
for nm, op, src in [("IntTDiv0y2", TDIV, "pir_test.cpp:837-845"), ("IntFDiv0y2", FDIV, "pir_test.cpp:847-855"),
                    ("IntEDiv0y2", EDIV, "pir_test.cpp:857-865")]:
    kat(nm + ".a", src, [[0, 1], [2, 2], [0, 0]], [[op, 2, 0, 1]], [[0, 1]], ua=False)
    kat(nm + ".b", src, [[1, 1], [2, 2], [0, 0]], [[op, 2, 0, 1]], [[1, 1]], ua=True)
kat("IntCDiv0y2", "pir_test.cpp:867-873", [[0, 1], [2, 2], [0, 0]], [[CDIV, 2, 0, 1]], [[0, 0]], ua=True)
for nm, op, src, ylb in [("IntEDiv1", EDIV, "pir_test.cpp:875-883", -20), ("IntCDiv1", CDIV, "pir_test.cpp:885-893", -20),
                         ("IntTDiv1", TDIV, "pir_test.cpp:895-903", -21), ("IntFDiv1", FDIV, "pir_test.cpp:905-913", -21)]:
    kat(nm, src, [[2, 10], [-25, 25], [-2, 3]], [[op, 0, 1, 2]], [[2, 10], [ylb, 25], [-2, 3]], ua=False)
kat("InfiniteDomain1.a", "pir_test.cpp:924-930", [TOP, [0, 1], [5, 5]], [[LEQ, 1, 0, 2]], [TOP, [0, 1]], ua=False)
kat("InfiniteDomain1.b", "pir_test.cpp:931-933", [TOP, [1, 1], [5, 5]], [[LEQ, 1, 0, 2]], [[NI, 5], [1, 1]], ua=True)
kat("InfiniteDomain2.b", "pir_test.cpp:943-945", [TOP, [0, 0], [5, 5]], [[LEQ, 1, 0, 2]], [[6, PI], [0, 0]], ua=True)
for nm, src, y, after in [("EqualConstraint1", "pir_test.cpp:520-524", [9, 10], [[9, 10], [9, 10]]),
                          ("EqualConstraint2", "pir_test.cpp:526-530", [1, 2], [[1, 2], [1, 2]]),
                          ("EqualConstraint3", "pir_test.cpp:532-536", [0, 11], [[1, 10], [1, 10]]),
                          ("EqualConstraint4", "pir_test.cpp:538-542", [5, 11], [[5, 10], [5, 10]])]:
    kat(nm, src, [[1, 10], y, [1, 1]], [[EQ, 2, 0, 1]], after, ua=False)
kat("NotEqualConstraint1", "pir_test.cpp:544-548", [[1, 10], [10, 10], [0, 0]], [[EQ, 2, 0, 1]], [[1, 9]], ua=True)
kat("NotEqualConstraint2", "pir_test.cpp:550-554", [[1, 10], [10, 10], [0, 0]], [[EQ, 2, 0, 1]],
    [[1, 9], [10, 10]], ua=True)
kat("NotEqualConstraint3", "pir_test.cpp:556-560", [[1, 10], [1, 1], [0, 0]], [[EQ, 2, 0, 1]],
    [[2, 10], [1, 1]], ua=True)
kat("Strict1", "pir_test.cpp:502-506", [[1, 10], [10, 10], [0, 0]], [[LEQ, 2, 0, 1]], None, bot=True)
kat("Strict2", "pir_test.cpp:508-512", [[1, 10], [10, 10], [0, 0]], [[LEQ, 2, 1, 0]], [[1, 9]], ua=True)

# ---- two and more records -----------------------------------------------------------------------------------------
# x + y + z <= k  :  t = x + y ; s = t + z ; s in [-inf, k]        vars: x y z t s
for nm, src, dom, k, after, bot, ua in [
    ("TopProp", "pir_test.cpp:322-329", [3, 10], 8, None, True, None),
    ("TernaryAdd2", "pir_test.cpp:332-339", [3, 10], 9, [[3, 3]] * 3, False, True),
    ("TernaryAdd3", "pir_test.cpp:342-349", [3, 10], 10, [[3, 4]] * 3, False, False),
    ("TernaryAdd4", "pir_test.cpp:352-359", [-2, 2], -5, [[-2, -1]] * 3, False, False),
]:
    kat(nm, src, [dom, dom, dom, TOP, [NI, k]], [[ADD, 3, 0, 1], [ADD, 4, 3, 2]], after, bot=bot, ua=ua)

# a*x + b*y + c*z <= 2 over 0/1 variables: vars x y z  A B C  ta tb tc  u s
def pseudo_boolean(a, b, c):
    store = [[0, 1], [0, 1], [0, 1], [a, a], [b, b], [c, c], TOP, TOP, TOP, TOP, [NI, 2]]
    recs = [[MUL, 6, 3, 0], [MUL, 7, 4, 1], [MUL, 8, 5, 2], [ADD, 9, 6, 7], [ADD, 10, 9, 8]]
    return store, recs


for nm, src, coefs, after in [
    ("PseudoBoolean1", "pir_test.cpp:362-369", (2, 1, 3), [[0, 1], [0, 1], [0, 0]]),
    ("PseudoBoolean2", "pir_test.cpp:372-384", (2, 5, 3), [[0, 1], [0, 0], [0, 0]]),
    ("PseudoBoolean3", "pir_test.cpp:387-394", (3, 5, 3), [[0, 0], [0, 0], [0, 0]]),
    ("PseudoBoolean4", "pir_test.cpp:397-404", (-1, 1, 3), [[0, 1], [0, 1], [0, 1]]),
]:
    st, rc = pseudo_boolean(*coefs)
    kat(nm, src, st, rc, after)

# -x <= y etc: vars x y ZERO ONE t ;  ZERO = t + x ; ONE = (t <= y)  /  ONE = (y <= t)  /  ZERO = (t <= y)
for nm, src, x, y, rel, after, bot in [
    ("NegationOp1", "pir_test.cpp:407-414", [-4, 3], 2, "le", [[-2, 3], [2, 2]], False),
    ("NegationOp2", "pir_test.cpp:417-423", [-4, 3], -2, "le", [[2, 3], [-2, -2]], False),
    ("NegationOp3", "pir_test.cpp:426-433", [0, 3], -2, "le", [[2, 3], [-2, -2]], False),
    ("NegationOp4", "pir_test.cpp:436-443", [-4, -3], 4, "le", [[-4, -3], [4, 4]], False),
    ("NegationOp5", "pir_test.cpp:446-453", [-4, 3], -2, "ge", [[-4, 2], [-2, -2]], False),
    ("NegationOp6", "pir_test.cpp:456-463", [-4, 3], 2, "gt", [[-4, -3], [2, 2]], False),
    ("NegationOp7", "pir_test.cpp:466-473", [-4, 3], 5, "ge", None, True),
]:
    rel_rec = {"le": [LEQ, 3, 4, 1], "ge": [LEQ, 3, 1, 4], "gt": [LEQ, 2, 4, 1]}[rel]
    kat(nm, src, [x, [y, y], [0, 0], [1, 1], TOP], [[ADD, 2, 4, 0], rel_rec], after, bot=bot)

# x - y <= k (x = t + y ; ONE = (t <= K)) and x - y >= k (ONE = (K <= t)): vars x y K ONE t
for nm, src, k, rel, after in [
    ("TemporalConstraint6", "pir_test.cpp:271-277", 5, "le", [[0, 10], [0, 10]]),
    ("TemporalConstraint7", "pir_test.cpp:280-286", -10, "le", [[0, 0], [10, 10]]),
    ("TemporalConstraint8", "pir_test.cpp:289-295", 5, "ge", [[5, 10], [0, 5]]),
    ("TemporalConstraint9", "pir_test.cpp:298-304", -5, "le", [[0, 5], [5, 10]]),
]:
    rel_rec = [LEQ, 3, 4, 2] if rel == "le" else [LEQ, 3, 2, 4]
    kat(nm, src, [[0, 10], [0, 10], [k, k], [1, 1], TOP], [[ADD, 0, 4, 1], rel_rec], after)

# x + y (rel) 5: vars x y FIVE ONE/ZERO t ; t = x + y
for nm, src, dom, rel, after, ua in [
    ("TemporalConstraint1", "pir_test.cpp:226-232", [0, 10], "le", [[0, 5], [0, 5]], False),
    ("TemporalConstraint2", "pir_test.cpp:235-241", [0, 10], "gt", [[0, 10], [0, 10]], False),
    ("TemporalConstraint3", "pir_test.cpp:244-250", [0, 3], "gt", [[3, 3], [3, 3]], True),
    ("TemporalConstraint4", "pir_test.cpp:253-259", [0, 3], "ge", [[2, 3], [2, 3]], False),
    ("TemporalConstraint5", "pir_test.cpp:262-268", [0, 4], "eq", [[1, 4], [1, 4]], False),
]:
    b = [0, 0] if rel == "gt" else [1, 1]
    rel_rec = {"le": [LEQ, 3, 4, 2], "gt": [LEQ, 3, 4, 2], "ge": [LEQ, 3, 2, 4], "eq": [EQ, 3, 4, 2]}[rel]
    kat(nm, src, [dom, dom, [5, 5], b, TOP], [[ADD, 4, 0, 1], rel_rec], after)

# b <=> (x - y <= k1 /\ y - x <= k2): vars x y b  K1 K2  t1 t2  b1 b2
def resource(x, y, b, k1, k2):
    store = [x, y, b, [k1, k1], [k2, k2], TOP, TOP, [0, 1], [0, 1]]
    recs = [[ADD, 0, 5, 1], [LEQ, 7, 5, 3], [ADD, 1, 6, 0], [LEQ, 8, 6, 4], [MIN, 2, 7, 8]]
    return store, recs


st, rc = resource([5, 10], [9, 15], [0, 1], 0, 2)
kat("ResourceConstraint1.a", "pir_test.cpp:476-484", st, rc, [[5, 10], [9, 15], [0, 1]])
st, rc = resource([5, 10], [9, 15], [1, 1], 0, 2)
kat("ResourceConstraint1.b", "pir_test.cpp:486-487", st, rc, [[7, 10], [9, 12], [1, 1]])
st, rc = resource([1, 2], [0, 2], [0, 1], 2, -1)
kat("ResourceConstraint2.a", "pir_test.cpp:490-497", st, rc, [[1, 2], [0, 2], [0, 1]])
st, rc = resource([1, 2], [0, 2], [0, 0], 2, -1)
kat("ResourceConstraint2.b", "pir_test.cpp:499-500", st, rc, [[1, 2], [1, 2], [0, 0]])

# (x == 5) xor (y == 5): vars x y FIVE ZERO b1 b2 ; b1 = (x == 5) ; b2 = (y == 5) ; ZERO = (b1 == b2)
xor_recs = [[EQ, 4, 0, 2], [EQ, 5, 1, 2], [EQ, 3, 4, 5]]
kat("XorConstraint1.a", "pir_test.cpp:584-589", [TOP, TOP, [5, 5], [0, 0], [0, 1], [0, 1]], xor_recs, [TOP, TOP])
kat("XorConstraint1.b", "pir_test.cpp:591-592", [[1, 1], TOP, [5, 5], [0, 0], [0, 1], [0, 1]], xor_recs,
    [[1, 1], [5, 5]])
kat("XorConstraint2.a", "pir_test.cpp:596-600", [[1, 5], [1, 5], [5, 5], [0, 0], [0, 1], [0, 1]], xor_recs,
    [[1, 5], [1, 5]])
kat("XorConstraint2.b", "pir_test.cpp:602-603", [[1, 5], [5, 5], [5, 5], [0, 0], [0, 1], [0, 1]], xor_recs,
    [[1, 4], [5, 5]])

# x in {1,3}: store layout of the reference's listing (pir_test.cpp:609-616): x C1 B0 C3 B1 y, then ONE
in_recs = [[EQ, 2, 0, 1], [EQ, 4, 0, 3], [MAX, 6, 2, 4]]
in_store = [[1, 3], [1, 1], [0, 1], [3, 3], [0, 1], [2, 3], [1, 1]]
kat("InConstraint1.a", "pir_test.cpp:606-617", in_store, in_recs, in_store[:6])
kat("InConstraint1.b", "pir_test.cpp:619-620", in_store, in_recs + [[EQ, 6, 0, 5]],
    [[3, 3], [1, 1], [0, 0], [3, 3], [1, 1], [3, 3]])

# min(b1,b2) = 1 then b1 = (x <= 5), b2 = (x >= 5): vars x b1 b2 ONE FIVE
kat("MinConstraint3.a", "pir_test.cpp:657-661", [TOP, [0, 1], [0, 1], [1, 1], [5, 5]], [[MIN, 3, 1, 2]],
    [TOP, [1, 1], [1, 1]])
kat("MinConstraint3.b", "pir_test.cpp:663-664", [TOP, [1, 1], [1, 1], [1, 1], [5, 5]],
    [[MIN, 3, 1, 2], [LEQ, 1, 0, 4]], [[NI, 5], [1, 1], [1, 1]])
kat("MinConstraint3.c", "pir_test.cpp:666-667", [[NI, 5], [1, 1], [1, 1], [1, 1], [5, 5]],
    [[MIN, 3, 1, 2], [LEQ, 1, 0, 4], [LEQ, 2, 4, 0]], [[5, 5], [1, 1], [1, 1]])
# max(b1,b2) = 0 then b1 = (x <= 5), b2 = (x >= 7): vars x b1 b2 ZERO FIVE SEVEN
kat("MaxConstraint3.a", "pir_test.cpp:698-702", [TOP, [0, 1], [0, 1], [0, 0], [5, 5], [7, 7]], [[MAX, 3, 1, 2]],
    [TOP, [0, 0], [0, 0]])
kat("MaxConstraint3.b", "pir_test.cpp:703-704", [TOP, [0, 0], [0, 0], [0, 0], [5, 5], [7, 7]],
    [[MAX, 3, 1, 2], [LEQ, 1, 0, 4]], [[6, PI], [0, 0], [0, 0]])
kat("MaxConstraint3.c", "pir_test.cpp:705-706", [[6, PI], [0, 0], [0, 0], [0, 0], [5, 5], [7, 7]],
    [[MAX, 3, 1, 2], [LEQ, 1, 0, 4], [LEQ, 2, 5, 0]], [[6, 6], [0, 0], [0, 0]])

# bool_clause([x1,x2],[y1,y2]) = x1 \/ x2 \/ not y1 \/ not y2: vars x1 x2 y1 y2 ZERO ONE n1 n2 t1 t2
clause_recs = [[EQ, 6, 2, 4], [EQ, 7, 3, 4], [MAX, 8, 0, 1], [MAX, 9, 6, 7], [MAX, 5, 8, 9]]


def clause(x1, x2, y1, y2):
    return [x1, x2, y1, y2, [0, 0], [1, 1], [0, 1], [0, 1], [0, 1], [0, 1]]


B = [0, 1]
kat("BooleanClause1", "pir_test.cpp:709-717", clause([1, 1], B, B, B), clause_recs, [[1, 1], B, B, B])
kat("BooleanClause2", "pir_test.cpp:719-727", clause(B, B, [0, 0], B), clause_recs, [B, B, [0, 0], B])
kat("BooleanClause3", "pir_test.cpp:729-741", clause([0, 0], [0, 0], [1, 1], B), clause_recs,
    [[0, 0], [0, 0], [1, 1], [0, 0]])
kat("BooleanClause4", "pir_test.cpp:743-755", clause([0, 0], B, [1, 1], [1, 1]), clause_recs,
    [[0, 0], [1, 1], [1, 1], [1, 1]])

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pir_kat.json")
    with open(out, "w") as f:
        json.dump(K, f, indent=0)
    print(f"wrote {len(K)} cases to {out}")
