"""Front-end of the PIR path: rewrite integer constraints into the ternary form `X = Y op Z` that PIR interprets
(lala-pc include/lala/pir.hpp:254-287).

The reference relies on lala-core's `ternarize` (lala/logic/ternarize.hpp, included at pir.hpp:15, un-vendored) to get
there from FlatZinc; this module is a restatement of that STEP, not of its text: the decomposition below is the
textbook one and its variable numbering is its own, so what is checked (tests/test_ternarize.py) is semantic — the
ternary network has exactly the solutions of the formula on the formula's variables — not a byte-for-byte match of
upstream's output. Constants become singleton variables (`ONE`, `ZERO`, `C5` in pir_test.cpp's hand-written networks),
every non-variable sub-term gets a fresh variable `t = a op b`, comparisons and connectives are reified into 0/1
variables (`r = (a <= b)`, `r = (a == b)`, and = min, or = max, not r = 1 - r, imply = (r1 <= r2), equiv = (r1 == r2)) and
a constraint asserted at top level uses the constant ONE (or ZERO for a negation) as its result variable, which is how
the reference's tests write `x <= y` (`LEQ(X = ONE, Y = x, Z = y)`).

Formulas are the nested tuples of pcflat.py:
  terms     ('var', v) ('const', k) ('neg', t) ('abs', t) ('add', a, b) ('sub', a, b) ('mul', a, b) ('sum', t1, ..., tn)
            ('min', a, b) ('max', a, b) ('tdiv' | 'fdiv' | 'cdiv' | 'ediv', a, b)
  formulas  ('le' | 'lt' | 'ge' | 'gt' | 'eq' | 'ne', a, b)  ('and' | 'or' | 'imply' | 'equiv', f, g)  ('not', f)
            ('lit', v) ('nlit', v)  ('true',) ('false',)
"""
import numpy as np

ADD, MUL, MIN, MAX, TDIV, FDIV, CDIV, EDIV, EQ, LEQ = 2, 4, 6, 7, 25, 27, 29, 31, 46, 48
INT_MIN, INT_MAX = -2**31, 2**31 - 1
_BIN = {"add": ADD, "mul": MUL, "min": MIN, "max": MAX, "tdiv": TDIV, "fdiv": FDIV, "cdiv": CDIV, "ediv": EDIV}
_FORMULA_OPS = ("le", "lt", "ge", "gt", "eq", "ne", "and", "or", "imply", "equiv", "not", "lit", "nlit", "true", "false")


class Ternarizer:
    """Accumulates records [op, x, y, z] (lala-core Sig codes, include/lpc.h) and the domains of the variables it adds.

    `store` is the [nvars, 2] domain array of the formula's own variables; `result()` returns the records and the
    extended store. Temporaries start at the interval hull of their definition (see `_hull`), constants at their value,
    reified results at [0, 1]."""

    def __init__(self, store):
        self.doms = [tuple(int(b) for b in d) for d in np.asarray(store).reshape(-1, 2)]
        self.n_original = len(self.doms)
        self.records = []
        self._consts = {}
        self._cse = {}

    # ---- variables -------------------------------------------------------------------------------------------------
    def _fresh(self, lb=INT_MIN, ub=INT_MAX):
        self.doms.append((lb, ub))
        return len(self.doms) - 1

    def const(self, k):
        if k not in self._consts:
            self._consts[k] = self._fresh(k, k)
        return self._consts[k]

    def _emit(self, op, x, y, z):
        self.records.append((op, x, y, z))

    def _hull(self, op, y, z):
        """Initial domain of t = y op z from the operands' domains. Not an optimisation only: a temporary left at top
        meets rules such as `x = 0 => y >= z.lb + 1` (pir.hpp:750-753) that turn an infinite bound into a finite one next
        to INT_MIN, after which `+` wraps around (the reference is undefined there, include/lpc.h) and the result
        depends on the evaluation order. An infinite operand bound keeps the corresponding side infinite."""
        (yl, yu), (zl, zu) = self.doms[y], self.doms[z]
        inf = lambda b: b in (INT_MIN, INT_MAX)
        if op == ADD:
            lo = INT_MIN if inf(yl) or inf(zl) else yl + zl
            hi = INT_MAX if inf(yu) or inf(zu) else yu + zu
        elif op == "sub":
            lo = INT_MIN if inf(yl) or inf(zu) else yl - zu
            hi = INT_MAX if inf(yu) or inf(zl) else yu - zl
        elif op == MUL:
            if any(inf(b) for b in (yl, yu, zl, zu)):
                return INT_MIN, INT_MAX
            p = [yl * zl, yl * zu, yu * zl, yu * zu]
            lo, hi = min(p), max(p)
        elif op == MIN:
            lo, hi = min(yl, zl), min(yu, zu)
        elif op == MAX:
            lo, hi = max(yl, zl), max(yu, zu)
        else:   # divisions: |y / z| <= |y|
            if inf(yl) or inf(yu):
                return INT_MIN, INT_MAX
            m = max(abs(yl), abs(yu))
            lo, hi = -m, m
        return max(lo, INT_MIN), min(hi, INT_MAX)

    def _def(self, op, y, z, boolean=False):
        """A fresh variable t with t = y op z; identical definitions share one variable."""
        key = (op, y, z)
        if key not in self._cse:
            t = self._fresh(0, 1) if boolean else self._fresh(*self._hull(op, y, z))
            self._emit(op, t, y, z)
            self._cse[key] = t
        return self._cse[key]

    # ---- terms -> the variable that holds their value ---------------------------------------------------------------
    def term(self, t):
        op = t[0]
        if op == "var":
            return int(t[1])
        if op == "const":
            return self.const(int(t[1]))
        if op in _BIN:
            return self._def(_BIN[op], self.term(t[1]), self.term(t[2]))
        if op == "sum":
            acc = self.term(t[1])
            for s in t[2:]:
                acc = self._def(ADD, acc, self.term(s))
            return acc
        if op == "sub":      # d = a - b  <=>  a = d + b
            a, b = self.term(t[1]), self.term(t[2])
            key = ("sub", a, b)
            if key not in self._cse:
                d = self._fresh(*self._hull("sub", a, b))
                self._emit(ADD, a, d, b)
                self._cse[key] = d
            return self._cse[key]
        if op == "neg":      # n = -a  <=>  0 = n + a
            return self.term(("sub", ("const", 0), t[1]))
        if op == "abs":      # |a| = max(a, -a)
            return self._def(MAX, self.term(t[1]), self.term(("neg", t[1])))
        if op in _FORMULA_OPS:   # a formula used as a 0/1 term (formula.hpp:1090-1102)
            return self.reify(t)
        raise ValueError(f"unknown term {op}")

    # ---- formulas -> a 0/1 variable equivalent to them ---------------------------------------------------------------
    def reify(self, f):
        op = f[0]
        if op == "true":
            return self.const(1)
        if op == "false":
            return self.const(0)
        if op == "lit":          # a Boolean variable is its own reification
            return int(f[1])
        if op == "nlit":
            return self.reify(("not", ("lit", f[1])))
        if op == "le":
            return self._def(LEQ, self.term(f[1]), self.term(f[2]), boolean=True)
        if op == "ge":
            return self.reify(("le", f[2], f[1]))
        if op == "gt":           # a > b  <=>  not (a <= b)
            return self.reify(("not", ("le", f[1], f[2])))
        if op == "lt":
            return self.reify(("not", ("le", f[2], f[1])))
        if op == "eq":
            return self._def(EQ, self.term(f[1]), self.term(f[2]), boolean=True)
        if op == "ne":
            return self.reify(("not", ("eq", f[1], f[2])))
        if op == "not":          # r = 1 - s  <=>  1 = r + s
            s = self.reify(f[1])
            key = ("not", s)
            if key not in self._cse:
                r = self._fresh(0, 1)
                self._emit(ADD, self.const(1), r, s)
                self._cse[key] = r
            return self._cse[key]
        if op == "and":
            return self._def(MIN, self.reify(f[1]), self.reify(f[2]), boolean=True)
        if op == "or":
            return self._def(MAX, self.reify(f[1]), self.reify(f[2]), boolean=True)
        if op == "imply":
            return self._def(LEQ, self.reify(f[1]), self.reify(f[2]), boolean=True)
        if op == "equiv":
            return self._def(EQ, self.reify(f[1]), self.reify(f[2]), boolean=True)
        raise ValueError(f"unknown formula {op}")

    # ---- constraints asserted at top level ---------------------------------------------------------------------------
    def tell(self, f, truth=True):
        """Assert that `f` holds (or, with truth=False, that it does not)."""
        op = f[0]
        def one():      # the result variable of a constraint that must hold ...
            return self.const(1 if truth else 0)

        def zero():     # ... and of one that must not (constants are created on first use)
            return self.const(0 if truth else 1)
        if op == "and" and truth:
            self.tell(f[1]); self.tell(f[2])
        elif op == "or" and not truth:
            self.tell(f[1], False); self.tell(f[2], False)
        elif op == "not":
            self.tell(f[1], not truth)
        elif op == "le":
            self._emit(LEQ, one(), self.term(f[1]), self.term(f[2]))
        elif op == "ge":
            self._emit(LEQ, one(), self.term(f[2]), self.term(f[1]))
        elif op == "gt":
            self._emit(LEQ, zero(), self.term(f[1]), self.term(f[2]))
        elif op == "lt":
            self._emit(LEQ, zero(), self.term(f[2]), self.term(f[1]))
        elif op == "eq":
            a, b = f[1], f[2]
            if truth and a[0] == "var" and b[0] in _BIN:      # x = y op z is already ternary
                self._emit(_BIN[b[0]], int(a[1]), self.term(b[1]), self.term(b[2]))
            elif truth and b[0] == "var" and a[0] in _BIN:
                self._emit(_BIN[a[0]], int(b[1]), self.term(a[1]), self.term(a[2]))
            else:
                self._emit(EQ, one(), self.term(a), self.term(b))
        elif op == "ne":
            self._emit(EQ, zero(), self.term(f[1]), self.term(f[2]))
        else:                                                    # literals and the other connectives: reify, then fix
            r = self.reify(f)
            self._emit(EQ, self.const(1), r, one())

    def result(self):
        recs = np.array(self.records, dtype=np.int32).reshape(-1, 4)
        store = np.array(self.doms, dtype=np.int64).astype(np.int32).reshape(-1, 2)
        return recs, store


def ternarize(formulas, store):
    """Formulas asserted together over `store` ([nvars, 2]) -> (records [n, 4], extended store, number of original vars)."""
    t = Ternarizer(store)
    for f in formulas:
        t.tell(f)
    recs, ext = t.result()
    return recs, ext, t.n_original
