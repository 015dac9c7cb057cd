"""Writes tests/golden/pc_kat.json: the reference's PC golden vectors (tests/pc_test.cpp) as formula trees.

The reference builds its propagators from FlatZinc through an un-vendored parser; here each case is the formula tree
the PC interpreter produces for it (pc.hpp:217-604): `int_ge(t, k)` is the flipped `k <= t` (pc.hpp:565),
`int_plus(x, y, k)` is `x + y = k`, `bool_clause([x1,x2],[y1,y2])` is the binarised disjunction
`x1 \\/ (x2 \\/ (not y1 \\/ not y2))` (pc.hpp:542-544), `b = (phi)` is the biconditional `b <=> phi`, a Boolean
variable used as a formula is a VariableLiteral. Expected intervals are the reference's own. `bot: true` means the
reference expects every listed variable to be bot and `is_bot()`.

Run:  python tests/golden/make_pc_kat.py
"""
import json
import os

NI, PI = -2**31, 2**31 - 1
TOP = [NI, PI]
K = []


def v(i):
    return ["var", i]


def c(k):
    return ["const", k]


def kat(name, source, store, props, after=None, bot=False, ua=None, changed=None):
    K.append(dict(name=name, source=source, store=store, props=props, after=after, bot=bot, ua=ua, changed=changed))


D10 = [0, 10]
kat("AddEquality", "pc_test.cpp:114-120", [D10, D10], [["eq", ["add", v(0), v(1)], c(5)]], [[0, 5], [0, 5]], ua=False, changed=True)
kat("TemporalConstraint1Flat", "pc_test.cpp:123-129", [D10, D10, [NI, 5]], [["eq", ["add", v(0), v(1)], v(2)]],
    [[0, 5], [0, 5], [0, 5]], ua=False, changed=True)
kat("TemporalConstraint1", "pc_test.cpp:132-138", [D10, D10], [["le", ["add", v(0), v(1)], c(5)]], [[0, 5], [0, 5]], ua=False, changed=True)
kat("TemporalConstraint2", "pc_test.cpp:141-147", [D10, D10], [["gt", ["add", v(0), v(1)], c(5)]], [D10, D10], ua=False, changed=False)
kat("TemporalConstraint3", "pc_test.cpp:150-156", [[0, 3], [0, 3]], [["gt", ["add", v(0), v(1)], c(5)]], [[3, 3], [3, 3]], ua=True, changed=True)
kat("TemporalConstraint4", "pc_test.cpp:159-165", [[0, 3], [0, 3]], [["le", c(5), ["add", v(0), v(1)]]], [[2, 3], [2, 3]], ua=False, changed=True)
kat("TemporalConstraint5", "pc_test.cpp:168-174", [[0, 4], [0, 4]], [["eq", ["add", v(0), v(1)], c(5)]], [[1, 4], [1, 4]], ua=False, changed=True)
kat("TemporalConstraint6", "pc_test.cpp:177-183", [D10, D10], [["le", ["sub", v(0), v(1)], c(5)]], [D10, D10], ua=False, changed=False)
kat("TemporalConstraint7", "pc_test.cpp:186-192", [D10, D10], [["le", ["sub", v(0), v(1)], c(-10)]], [[0, 0], [10, 10]], ua=True, changed=True)
kat("TemporalConstraint8", "pc_test.cpp:195-201", [D10, D10], [["le", c(5), ["sub", v(0), v(1)]]], [[5, 10], [0, 5]], ua=False, changed=True)
kat("TemporalConstraint9", "pc_test.cpp:204-210", [D10, D10], [["le", ["sub", v(0), v(1)], c(-5)]], [[0, 5], [5, 10]], ua=False, changed=True)
kat("TemporalConstraint10", "pc_test.cpp:213-219", [D10, D10], [["le", v(0), ["add", c(-5), v(1)]]], [[0, 5], [5, 10]], ua=False, changed=True)
for nm, src, dom, k, after, bot, ua in [
    ("TopProp", "pc_test.cpp:222-230", [3, 10], 8, None, True, None),
    ("TernaryAdd2", "pc_test.cpp:233-240", [3, 10], 9, [[3, 3]] * 3, False, True),
    ("TernaryAdd3", "pc_test.cpp:243-250", [3, 10], 10, [[3, 4]] * 3, False, False),
    ("TernaryAdd4", "pc_test.cpp:253-260", [-2, 2], -5, [[-2, -1]] * 3, False, False),
]:
    kat(nm, src, [dom, dom, dom], [["le", ["add", ["add", v(0), v(1)], v(2)], c(k)]], after, bot=bot, ua=ua, changed=True)
    # the same constraint as ONE n-ary sum (the shape config 3 uses); same expected result
    kat(nm + ".nary", src, [dom, dom, dom], [["le", ["sum", v(0), v(1), v(2)], c(k)]], after, bot=bot, ua=ua, changed=True)
B = [0, 1]
for nm, src, terms, after, ua, chg in [
    ("PseudoBoolean1", "pc_test.cpp:263-270", (["mul", c(2), v(0)], v(1), ["mul", c(3), v(2)]), [B, B, [0, 0]], False, True),
    ("PseudoBoolean2", "pc_test.cpp:273-280", (["mul", c(2), v(0)], ["mul", c(5), v(1)], ["mul", c(3), v(2)]), [B, [0, 0], [0, 0]], True, True),
    ("PseudoBoolean3", "pc_test.cpp:283-290", (["mul", c(3), v(0)], ["mul", c(5), v(1)], ["mul", c(3), v(2)]), [[0, 0]] * 3, True, True),
    ("PseudoBoolean4", "pc_test.cpp:293-300", (["neg", v(0)], v(1), ["mul", c(3), v(2)]), [B, B, B], False, False),
]:
    kat(nm, src, [B, B, B], [["le", ["add", ["add", terms[0], terms[1]], terms[2]], c(2)]], after, ua=ua, changed=chg)
    kat(nm + ".nary", src, [B, B, B], [["le", ["sum", terms[0], terms[1], terms[2]], c(2)]], after, ua=ua, changed=chg)
for nm, src, dom, prop, after, bot, ua in [
    ("NegationOp1", "pc_test.cpp:303-308", [-4, 3], ["le", ["neg", v(0)], c(2)], [[-2, 3]], False, True),
    ("NegationOp2", "pc_test.cpp:311-316", [-4, 3], ["le", ["neg", v(0)], c(-2)], [[2, 3]], False, True),
    ("NegationOp3", "pc_test.cpp:319-324", [0, 3], ["le", ["neg", v(0)], c(-2)], [[2, 3]], False, True),
    ("NegationOp4", "pc_test.cpp:327-332", [-4, -3], ["le", ["neg", v(0)], c(4)], [[-4, -3]], False, True),
    ("NegationOp5", "pc_test.cpp:335-340", [-4, 3], ["le", c(-2), ["neg", v(0)]], [[-4, 2]], False, True),
    ("NegationOp6", "pc_test.cpp:343-348", [-4, 3], ["gt", ["neg", v(0)], c(2)], [[-4, -3]], False, True),
    ("NegationOp7", "pc_test.cpp:351-357", [-4, 3], ["le", c(5), ["neg", v(0)]], None, True, None),
]:
    kat(nm, src, [dom], [prop], after, bot=bot, ua=ua)


def resource(k1, k2):
    return ["equiv", ["lit", 2], ["and", ["le", ["sub", v(0), v(1)], c(k1)], ["le", ["sub", v(1), v(0)], c(k2)]]]


kat("ResourceConstraint1.a", "pc_test.cpp:360-368", [[5, 10], [9, 15], B], [resource(0, 2)], [[5, 10], [9, 15], B], ua=False, changed=False)
kat("ResourceConstraint1.b", "pc_test.cpp:370-371", [[5, 10], [9, 15], [1, 1]], [resource(0, 2)], [[7, 10], [9, 12], [1, 1]], ua=False, changed=True)
kat("ResourceConstraint2.a", "pc_test.cpp:374-381", [[1, 2], [0, 2], B], [resource(2, -1)], [[1, 2], [0, 2], B], ua=False, changed=False)
kat("ResourceConstraint2.b", "pc_test.cpp:383-384", [[1, 2], [0, 2], [0, 0]], [resource(2, -1)], [[1, 2], [1, 2], [0, 0]], ua=False, changed=True)
kat("NotEqualConstraint1", "pc_test.cpp:386-390", [[1, 10]], [["ne", v(0), c(10)]], [[1, 9]], ua=True, changed=True)
kat("NotEqualConstraint2", "pc_test.cpp:392-396", [[1, 10], [10, 10]], [["ne", v(0), v(1)]], [[1, 9], [10, 10]], ua=True, changed=True)
kat("NotEqualConstraint3", "pc_test.cpp:398-402", [[1, 10]], [["ne", v(0), c(10)]], [[1, 9]], ua=True, changed=True)


def clause():
    return ["or", ["lit", 0], ["or", ["lit", 1], ["or", ["nlit", 2], ["nlit", 3]]]]


kat("BooleanClause1", "pc_test.cpp:564-572", [[1, 1], B, B, B], [clause()], [[1, 1], B, B, B], ua=True, changed=False)
kat("BooleanClause2", "pc_test.cpp:574-582", [B, B, [0, 0], B], [clause()], [B, B, [0, 0], B], ua=True, changed=False)
kat("BooleanClause3.c", "pc_test.cpp:584-596", [[0, 0], [0, 0], [1, 1], B], [clause()], [[0, 0], [0, 0], [1, 1], [0, 0]], ua=True, changed=True)
kat("BooleanClause4.c", "pc_test.cpp:598-610", [[0, 0], B, [1, 1], [1, 1]], [clause()], [[0, 0], [1, 1], [1, 1], [1, 1]], ua=True, changed=True)
kat("BooleanClause3.a", "pc_test.cpp:591-592", [[0, 0], B, B, B], [clause()], [[0, 0], B, B, B], ua=False, changed=False)
kat("IntAbs1", "pc_test.cpp:700-707", [[-15, 5], [-10, 10]], [["eq", ["abs", v(0)], v(1)]], [[-10, 5], [0, 10]], ua=False, changed=True)
kat("InfiniteDomain1.a", "pc_test.cpp:766-772", [TOP, B], [["equiv", ["lit", 1], ["le", v(0), c(5)]]], [TOP, B], ua=False, changed=False)
kat("InfiniteDomain1.b", "pc_test.cpp:773-775", [TOP, [1, 1]], [["equiv", ["lit", 1], ["le", v(0), c(5)]]], [[NI, 5], [1, 1]], ua=True, changed=True)
kat("InfiniteDomain2.b", "pc_test.cpp:785-787", [TOP, [0, 0]], [["equiv", ["lit", 1], ["le", v(0), c(5)]]], [[6, PI], [0, 0]], ua=True, changed=True)
# x = 5 xor y = 5  ==  (x = 5) <=> not (y = 5)  ==  (x = 5) <=> (y != 5)
xor = ["equiv", ["eq", v(0), c(5)], ["ne", v(1), c(5)]]
kat("XorConstraint1.b", "pc_test.cpp:434-435", [[1, 1], TOP], [xor], [[1, 1], [5, 5]], ua=True, changed=True)
kat("XorConstraint2.b", "pc_test.cpp:445-446", [[1, 5], [5, 5]], [xor], [[1, 4], [5, 5]], ua=True, changed=True)
# bool_xor builds an ExclusiveDisjunction node (pc.hpp:391-395, formula.hpp:518-587): the same two cases on that node
xor2 = ["xor", ["eq", v(0), c(5)], ["eq", v(1), c(5)]]
kat("XorConstraint1.a.xor", "pc_test.cpp:429-432", [TOP, TOP], [xor2], [TOP, TOP], ua=False, changed=False)
kat("XorConstraint1.b.xor", "pc_test.cpp:434-435", [[1, 1], TOP], [xor2], [[1, 1], [5, 5]], ua=True, changed=True)
kat("XorConstraint2.a.xor", "pc_test.cpp:440-443", [[1, 5], [1, 5]], [xor2], [[1, 5], [1, 5]], ua=False, changed=False)
kat("XorConstraint2.b.xor", "pc_test.cpp:445-446", [[1, 5], [5, 5]], [xor2], [[1, 4], [5, 5]], ua=True, changed=True)
# `var {1, 3}: x` typed into PC is the disjunction x = 1 \/ x = 3 (one propagator, pc_test.cpp:470-472); then x = y
inx = ["or", ["eq", v(0), c(1)], ["eq", v(0), c(3)]]
kat("InConstraint1.a", "pc_test.cpp:470-472", [[1, 3], [2, 3]], [inx], [[1, 3], [2, 3]], ua=False, changed=False)
kat("InConstraint1.b", "pc_test.cpp:474-475", [[1, 3], [2, 3]], [inx, ["eq", v(0), v(1)]], [[3, 3], [3, 3]], ua=True, changed=True)
# int_min(x, y, z) / int_max(x, y, z): Equality(Binary<GroupMinMax>(x, y), z) (pc.hpp:241-242, terms.hpp:301-331)
mn, mx = ["eq", ["min", v(0), v(1)], v(2)], ["eq", ["max", v(0), v(1)], v(2)]
kat("MinConstraint1.a", "pc_test.cpp:479-483", [[0, 4], [2, 5], [0, 10]], [mn], [[0, 4], [2, 5], [0, 4]], ua=False, changed=True)
kat("MinConstraint1.b", "pc_test.cpp:485-486", [[0, 4], [2, 5], [0, 3]], [mn], [[0, 4], [2, 5], [0, 3]], ua=False, changed=False)
kat("MinConstraint1.c", "pc_test.cpp:488-489", [[0, 1], [2, 5], [0, 3]], [mn], [[0, 1], [2, 5], [0, 1]], ua=False, changed=True)
kat("MinConstraint1.d", "pc_test.cpp:491-492", [[0, 0], [2, 5], [0, 1]], [mn], [[0, 0], [2, 5], [0, 0]], ua=True, changed=True)
kat("MinConstraint2.c", "pc_test.cpp:505-506", [[4, 4], [2, 5], [0, 3]], [mn], [[4, 4], [2, 3], [2, 3]], ua=False, changed=True)
mn3 = ["eq", ["min", v(1), v(2)], c(1)]
b1le, b2ge = ["equiv", ["lit", 1], ["le", v(0), c(5)]], ["equiv", ["lit", 2], ["le", c(5), v(0)]]
kat("MinConstraint3.a", "pc_test.cpp:510-514", [TOP, B, B], [mn3], [TOP, [1, 1], [1, 1]], ua=True, changed=True)
kat("MinConstraint3.b", "pc_test.cpp:516-517", [TOP, [1, 1], [1, 1]], [mn3, b1le], [[NI, 5], [1, 1], [1, 1]], ua=True, changed=True)
kat("MinConstraint3.c", "pc_test.cpp:519-520", [[NI, 5], [1, 1], [1, 1]], [mn3, b1le, b2ge], [[5, 5], [1, 1], [1, 1]], ua=True, changed=True)
kat("MaxConstraint1.a", "pc_test.cpp:524-528", [[0, 4], [2, 5], [0, 10]], [mx], [[0, 4], [2, 5], [2, 5]], ua=False, changed=True)
kat("MaxConstraint1.b", "pc_test.cpp:530-531", [[0, 4], [2, 5], [2, 3]], [mx], [[0, 3], [2, 3], [2, 3]], ua=False, changed=True)
kat("MaxConstraint1.c", "pc_test.cpp:533-534", [[0, 1], [2, 3], [2, 3]], [mx], [[0, 1], [2, 3], [2, 3]], ua=False, changed=False)
kat("MaxConstraint1.d", "pc_test.cpp:536-537", [[0, 1], [2, 2], [2, 3]], [mx], [[0, 1], [2, 2], [2, 2]], ua=True, changed=True)
kat("MaxConstraint2.b", "pc_test.cpp:547-548", [[0, 4], [2, 5], [5, 5]], [mx], [[0, 4], [5, 5], [5, 5]], ua=True, changed=True)
mx3 = ["eq", ["max", v(1), v(2)], c(0)]
b2ge7 = ["equiv", ["lit", 2], ["le", c(7), v(0)]]
kat("MaxConstraint3.a", "pc_test.cpp:551-555", [TOP, B, B], [mx3], [TOP, [0, 0], [0, 0]], ua=True, changed=True)
kat("MaxConstraint3.b", "pc_test.cpp:557-558", [TOP, [0, 0], [0, 0]], [mx3, b1le], [[6, PI], [0, 0], [0, 0]], ua=True, changed=True)
kat("MaxConstraint3.c", "pc_test.cpp:560-561", [[6, PI], [0, 0], [0, 0]], [mx3, b1le, b2ge7], [[6, 6], [0, 0], [0, 0]], ua=True, changed=True)
# int_times(x, y, z): Equality(Binary<GroupMul<EDIV>>(x, y), z) (pc.hpp:236, terms.hpp:231-262)
tm = ["eq", ["mul", v(0), v(1)], v(2)]
kat("IntTimes1.a", "pc_test.cpp:612-619", [B, B, B], [tm], [B, B, B], ua=False, changed=False)
kat("IntTimes1.b", "pc_test.cpp:620-621", [[1, 1], B, B], [tm], [[1, 1], B, B], ua=False, changed=False)
kat("IntTimes1.c", "pc_test.cpp:622-623", [[1, 1], [1, 1], B], [tm], [[1, 1], [1, 1], [1, 1]], ua=True, changed=True)
kat("IntTimes2.b", "pc_test.cpp:634-635", [B, B, [1, 1]], [tm], [[1, 1], [1, 1], [1, 1]], ua=True, changed=True)
kat("IntTimes3", "pc_test.cpp:638-646", [[0, 0], B, B], [tm], [[0, 0], B, [0, 0]], ua=True, changed=True)
kat("IntTimes4", "pc_test.cpp:648-656", [B, [0, 0], B], [tm], [B, [0, 0], [0, 0]], ua=True, changed=True)
kat("IntTimes5", "pc_test.cpp:658-666", [[1, 2], B, [0, 0]], [tm], [[1, 2], [0, 0], [0, 0]], ua=True, changed=True)
kat("IntTimes6", "pc_test.cpp:668-676", [B, [1, 2], [0, 0]], [tm], [[0, 0], [1, 2], [0, 0]], ua=True, changed=True)
# int_div(x, y, z): Equality(Binary<GroupDiv<TDIV>>(x, y), z) (pc.hpp:237, terms.hpp:264-299)
dv = ["eq", ["tdiv", v(0), v(1)], v(2)]
kat("IntDiv1.a", "pc_test.cpp:678-685", [B, B, B], [dv], [B, [1, 1], B], ua=False, changed=True)
kat("IntDiv1.b", "pc_test.cpp:686-687", [[1, 1], [1, 1], B], [dv], [[1, 1], [1, 1], [1, 1]], ua=True, changed=True)
dv2 = ["eq", ["tdiv", v(0), c(2)], c(0)]
kat("IntDiv2.a", "pc_test.cpp:690-695", [B], [dv2], [B], ua=True, changed=False)
kat("IntDiv2.b", "pc_test.cpp:696-697", [[1, 1]], [dv2], [[1, 1]], ua=True, changed=False)
# nbool_equiv with both sides typed into the store: Biconditional of two AbstractElements (formula.hpp:14-77)
ae1 = ["equiv", ["ae", "ge", 0, 5], ["ae", "le", 1, 5]]
kat("AbstractElement1.a", "pc_test.cpp:714-720", [D10, D10], [ae1], [D10, D10], ua=False, changed=False)
kat("AbstractElement1.b", "pc_test.cpp:721-722", [[5, 5], D10], [ae1], [[5, 5], [0, 5]], ua=True, changed=True)
kat("AbstractElement2.b", "pc_test.cpp:733-734", [[4, 4], D10], [ae1], [[4, 4], [6, 10]], ua=True, changed=True)
ae3 = ["equiv", ["ae", "eq", 0, 5], ["ae", "eq", 1, 5]]
kat("AbstractElement3.a", "pc_test.cpp:738-744", [D10, D10], [ae3], [D10, D10], ua=False, changed=False)
kat("AbstractElement3.b", "pc_test.cpp:746-747", [D10, [6, 6]], [ae3], [D10, [6, 6]], ua=False, changed=False)
kat("AbstractElement4.b", "pc_test.cpp:760-761", [D10, [4, 4]], [ae3], [D10, [4, 4]], ua=False, changed=False)
# array_int_element(b, [10, 11, 12], c): three propagators (pc_test.cpp:404-418). The decomposition happens in the
# un-vendored FlatZinc front-end; ASSUMED here: one implication per element, b = i => c = a[i] (Implication,
# formula.hpp:453-516) - it reproduces all three expected steps.
elem = [["imply", ["eq", v(0), c(i + 1)], ["eq", v(1), c(10 + i)]] for i in range(3)]
kat("ElementConstraint1.a", "pc_test.cpp:404-411", [[1, 3], [10, 12]], elem, [[1, 3], [10, 12]], ua=False, changed=False)
kat("ElementConstraint1.b", "pc_test.cpp:413-414", [[1, 3], [10, 11]], elem, [[1, 2], [10, 11]], ua=False, changed=True)
kat("ElementConstraint1.c", "pc_test.cpp:416-417", [[1, 2], [11, 11]], elem, [[2, 2], [11, 11]], ua=True, changed=True)

TERM_KATS = [
    dict(name="TermTest.AddTermBinary", source="pc_test.cpp:31-47", store=[D10, D10], term=["add", v(0), v(1)],
         project=[0, 20], embed=[NI, 5], changed=True, project_after=[0, 10]),
    dict(name="TermTest.AddTermNary", source="pc_test.cpp:49-67", store=[D10, D10, D10], term=["sum", v(0), v(1), v(2)],
         project=[0, 30], embed=[NI, 5], changed=True, project_after=[0, 15]),
]

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pc_kat.json")
    with open(out, "w") as f:
        json.dump(dict(props=K, terms=TERM_KATS), f, indent=0)
    print(f"wrote {len(K)} + {len(TERM_KATS)} cases to {out}")
