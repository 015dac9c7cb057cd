"""Writes tests/golden/pc_bitset_kat.json: the reference's bitset-domain golden vectors (tests/pc_bitset_test.cpp) as
formula trees over NBitset<64> cells.

Each case is the formula tree the PC interpreter builds (see make_pc_kat.py) and the NBit values the reference test
expects before / after `GaussSeidelIteration::fixpoint(bpc.num_deductions(), bpc.deduce)` (pc_bitset_test.cpp:44-62).
Cells are written symbolically, ["r", lb, ub] = NBit(lb, ub) and ["s", [v...]] = NBit::from_set({v...}), plus the
uint64 they stand for under the layout pinned by `var -15..5` == NBit(-1, 5) (pc_bitset_test.cpp:152-156): bit 0 =
"<= -1", bit i = value i - 1, bit 63 = ">= 62". Unary constraints (`int_ne(x, 10)`, `int_eq(x[1], true)`) are absorbed
by the store in the reference (0 deductions), so they appear here as the store the test asserts before the fixpoint.

Run:  python tests/golden/make_pc_bitset_kat.py
"""
import json
import os

K = []


def cell(spec):
    def rng(l, u):
        if l > u:
            return 0
        frm = 0 if l < 0 else (63 if l >= 62 else l + 1)
        to = 0 if u < 0 else (63 if u >= 62 else u + 1)
        return sum(1 << i for i in range(frm, to + 1))
    if spec[0] == "r":
        return rng(spec[1], spec[2])
    b = 0
    for v in spec[1]:
        b |= rng(v, v)
    return b


def kat(name, source, before, props, after=None, ua=None, changed=None):
    after = before if after is None else after
    K.append(dict(name=name, source=source, before=before, after=after, props=props, ua=ua, changed=changed,
                  before_bits=[cell(c) for c in before], after_bits=[cell(c) for c in after]))


def v(i):
    return ["var", i]


def r(l, u):
    return ["r", l, u]


B, T, F = r(0, 1), r(1, 1), r(0, 0)
not4 = ["s", [1, 2, 3, 5, 6, 7, 8, 9, 10]]
clause = ["or", ["lit", 0], ["or", ["lit", 1], ["or", ["nlit", 2], ["nlit", 3]]]]

kat("NotEqualConstraint1", "pc_bitset_test.cpp:68-72", [r(1, 9)], [], ua=True, changed=False)
kat("NotEqualConstraint2", "pc_bitset_test.cpp:74-78", [r(1, 10), r(10, 10)], [["ne", v(0), v(1)]], [r(1, 9), r(10, 10)], ua=True, changed=True)
kat("NotEqualConstraint3", "pc_bitset_test.cpp:80-84", [not4], [], ua=True, changed=False)
kat("NotEqualConstraint4", "pc_bitset_test.cpp:86-90", [r(1, 10), r(4, 4)], [["ne", v(0), v(1)]], [not4, r(4, 4)], ua=True, changed=True)
kat("InConstraint1.a", "pc_bitset_test.cpp:93-96", [["s", [1, 3]], r(2, 3)], [], ua=True, changed=False)
kat("InConstraint1.b", "pc_bitset_test.cpp:98-99", [["s", [1, 3]], r(2, 3)], [["eq", v(0), v(1)]], [r(3, 3), r(3, 3)], ua=True, changed=True)
kat("BooleanClause1.a", "pc_bitset_test.cpp:102-107", [B, B, B, B], [clause], ua=False, changed=False)
kat("BooleanClause1.b", "pc_bitset_test.cpp:108-109", [T, B, B, B], [clause], ua=True, changed=False)
kat("BooleanClause2.b", "pc_bitset_test.cpp:118-119", [B, B, F, B], [clause], ua=True, changed=False)
kat("BooleanClause3.b", "pc_bitset_test.cpp:128-129", [F, B, B, B], [clause], ua=False, changed=False)
kat("BooleanClause3.c", "pc_bitset_test.cpp:130-131", [F, F, B, B], [clause], ua=False, changed=False)
kat("BooleanClause3.d", "pc_bitset_test.cpp:132-133", [F, F, T, B], [clause], [F, F, T, F], ua=True, changed=True)
kat("BooleanClause4.c", "pc_bitset_test.cpp:144-145", [F, B, T, B], [clause], ua=False, changed=False)
kat("BooleanClause4.d", "pc_bitset_test.cpp:146-147", [F, B, T, T], [clause], [F, T, T, T], ua=True, changed=True)
kat("IntAbs1", "pc_bitset_test.cpp:150-157", [r(-1, 5), r(-1, 10)], [["eq", ["abs", v(0)], v(1)]], [r(-1, 5), r(0, 10)], ua=False, changed=True)

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pc_bitset_kat.json")
    with open(out, "w") as f:
        json.dump(dict(props=K), f, indent=0)
    print(f"wrote {len(K)} cases to {out}")
