"""Host-side flattening of PC formulas into the device encoding of include/lpc_pc.h — the counterpart of
PC::interpret_formula / interpret_term (lala-pc include/lala/pc.hpp:217-604) for the in-scope shapes.

Formulas are nested tuples (a plain Python stand-in for lala-core's TFormula):
  terms     ('var', v) ('const', k) ('neg', ('var', v)) ('mul', ('const', c), ('var', v)) ('add', t1, t2)
            ('sub', t1, ('var', v)) ('sum', t1, ..., tn >= 3) ('abs', ('var', v))
  formulas  ('le', term, ('const', k))  ('le', ('const', k), term)  ('gt', term, ('const', k))
            ('eq', term, ('const', k))  ('eq', sum-term, ('var', z))
            ('equiv', ('lit', b), ('le', term, ('const', k)))
            ('eq', ('var', x), ('var', y))  ('ne', ('var', x), ('var', y) | ('const', k))
            ('or', lit, ('or', lit, ...)) with lit = ('lit', v) | ('nlit', v)   ('eq', ('abs', ('var', x)), ('var', y))
Any other formula over these node types - and ('min' | 'max' | 'mul' | 'tdiv' | 'fdiv' | 'cdiv' | 'ediv', t1, t2) terms,
('prod', t1, ..., tn) products, ('and' | 'or' | 'equiv' | 'imply' | 'xor', f, g) connectives, ('true',) / ('false',), ('ae', op, var, k) store
elements, comparisons between two non-constant terms - keeps its tree: `flatten` encodes it as an
LPC_PC_TREE propagator (the prefix stream of include/lpc_pc.h in the propagator's term slots), which the device walks
node by node (csrc/pc_tree.cuh). `flatten(..., tree=False)` raises `Unsupported` for those instead (the reference's
"shape of this formula is not supported" interpretation error); trees deeper than the device limits always do.
"""
import numpy as np

PC_LIN_LE, PC_REIF_LIN_LE, PC_EQ, PC_NEQ, PC_CLAUSE, PC_ABS_EQ = 1, 2, 3, 4, 5, 6
PC_LIN_GE, PC_LIN_GT, PC_LIN_EQ, PC_LIN_EQ_VAR, PC_TREE = 7, 8, 9, 10, 11
TREE_TERM_DEPTH, TREE_FORM_DEPTH = 8, 6   # csrc/pc_tree.cuh

_T = {"const": 1, "var": 2, "neg": 3, "abs": 4, "add": 5, "sub": 6, "mul": 7, "sum": 8, "min": 9, "max": 10,
      "tdiv": 11, "fdiv": 12, "cdiv": 13, "ediv": 14, "prod": 15}
_F = {"lit": 20, "nlit": 21, "le": 22, "gt": 23, "eq": 24, "ne": 25, "and": 26, "or": 27, "equiv": 28, "imply": 29, "xor": 30,
      "ae": 31, "true": 32, "false": 33}
_AE = {"le": 0, "ge": 1, "eq": 2, "ne": 3}


class Unsupported(ValueError):
    pass


def _linear_terms(t, nb=False):
    """A term that is a variable, coef * variable, or an n-ary sum of those -> [(coef, var)].
    nb: the table is for NBitset stores. There -x and (-1) * x are different sets (the product goes through the constant's
    image in the universe, which folds every negative value into one bit), so Unary<Neg> and Binary<Sub> are not folded
    into a coefficient of -1: such formulas keep their tree."""
    op = t[0]
    if nb and op in ("neg", "sub"):
        raise Unsupported("Unary<Neg> / Binary<Sub> keep their tree over NBitset stores")
    if op == "var":
        return [(1, int(t[1]))]
    if op == "neg" and t[1][0] == "var":   # Unary<Neg>(Variable): the bounds of -1 * x (terms.hpp:87-102)
        return [(-1, int(t[1][1]))]
    if op == "mul" and t[1][0] == "const" and t[2][0] == "var" and int(t[1][1]) != 0:
        return [(int(t[1][1]), int(t[2][1]))]
    if op == "sub":   # Binary<GroupSub>(leaf, Variable): same residuals as leaf + (-1) * y (terms.hpp:209-229)
        a = _linear_terms(t[1], nb)
        if len(a) != 1 or t[1][0] in ("add", "sum", "sub") or t[2][0] != "var":
            raise Unsupported("only leaf - variable differences are flattened (a scaled right operand rounds differently)")
        return a + [(-1, int(t[2][1]))]
    if op == "add":   # Binary<GroupAdd> of two leaf terms
        a, b = _linear_terms(t[1], nb), _linear_terms(t[2], nb)
        if len(a) != 1 or len(b) != 1 or t[1][0] in ("add", "sum", "sub") or t[2][0] in ("add", "sum", "sub"):
            raise Unsupported("nested binary sums are not flattened (their residuals differ from a flat sum)")
        return a + b
    if op == "sum":
        if len(t) - 1 < 3:
            raise Unsupported("an n-ary sum has at least three operands (pc.hpp:277-296)")
        out = []
        for s in t[1:]:
            if s[0] in ("sum", "add", "sub"):
                raise Unsupported("nested n-ary sums are not flattened (their residuals differ from a flat sum)")
            out += _linear_terms(s, nb)
        return out
    raise Unsupported(f"term {op} is not linear")


def _lin_le(f, nb=False):
    if f[0] == "le" and f[2][0] == "const":
        return _linear_terms(f[1], nb), int(f[2][1])
    raise Unsupported("not a linear inequality with a constant right-hand side")


def _clause_literals(f):
    lits = []
    while True:
        if f[0] == "or" and f[1][0] in ("lit", "nlit"):
            lits.append((1 if f[1][0] == "lit" else -1, int(f[1][1])))
            f = f[2]
        elif f[0] in ("lit", "nlit"):
            lits.append((1 if f[0] == "lit" else -1, int(f[1])))
            return lits
        else:
            raise Unsupported("not a right-nested clause of literals")


def flatten_one(f, nb=False):
    """-> (kind, [(coef, var)], rhs, bvar); nb: for NBitset stores (see _linear_terms)"""
    op = f[0]
    if op == "le" and f[1][0] == "const" and f[2][0] != "const":   # k <= term
        return PC_LIN_GE, _linear_terms(f[2], nb), int(f[1][1]), -1
    if op == "le":
        terms, k = _lin_le(f, nb)
        return PC_LIN_LE, terms, k, -1
    if op == "gt" and f[2][0] == "const":
        return PC_LIN_GT, _linear_terms(f[1], nb), int(f[2][1]), -1
    if op == "eq" and f[2][0] == "const" and f[1][0] != "abs":
        return PC_LIN_EQ, _linear_terms(f[1], nb), int(f[2][1]), -1
    if op == "eq" and f[2][0] == "var" and f[1][0] in ("add", "sub", "sum", "mul", "neg"):
        return PC_LIN_EQ_VAR, _linear_terms(f[1], nb), 0, int(f[2][1])
    if op == "equiv" and f[1][0] == "lit" and f[2][0] == "le":
        terms, k = _lin_le(f[2], nb)
        return PC_REIF_LIN_LE, terms, k, int(f[1][1])
    if op == "eq" and f[1][0] == "var" and f[2][0] == "var":
        return PC_EQ, [(1, int(f[1][1])), (1, int(f[2][1]))], 0, -1
    if op == "eq" and f[1][0] == "abs" and f[1][1][0] == "var" and f[2][0] == "var":
        return PC_ABS_EQ, [(1, int(f[1][1][1])), (1, int(f[2][1]))], 0, -1
    if op == "ne" and f[1][0] == "var" and f[2][0] == "var":
        return PC_NEQ, [(1, int(f[1][1])), (1, int(f[2][1]))], 0, -1
    if op == "ne" and f[1][0] == "var" and f[2][0] == "const":
        return PC_NEQ, [(1, int(f[1][1]))], int(f[2][1]), -1
    if op in ("or", "lit", "nlit"):   # a lone VariableLiteral is a one-literal clause
        return PC_CLAUSE, _clause_literals(f), 0, -1
    raise Unsupported(f"formula {op} has no flat kind")


def _encode_term(t, out):
    """Prefix words of a term; returns its height."""
    op = t[0]
    if op not in _T:
        raise Unsupported(f"term {op} has no device rule")
    if op in ("const", "var"):
        out += [_T[op], int(t[1])]
        return 1
    if op in ("neg", "abs"):
        out.append(_T[op])
        return 1 + _encode_term(t[1], out)
    if op in ("sum", "prod"):
        if len(t) - 1 < 2:
            raise Unsupported("an n-ary sum or product has at least two operands")
        out += [_T[op], len(t) - 1]
        return 1 + max(_encode_term(s, out) for s in t[1:])
    out.append(_T[op])
    return 1 + max(_encode_term(t[1], out), _encode_term(t[2], out))


def _encode_formula(f, out):
    """Prefix words of a formula; returns the height of its connective nest (a comparison or literal is 1)."""
    op = f[0]
    if op not in _F:
        raise Unsupported(f"formula {op} has no device rule")
    if op in ("lit", "nlit"):
        out += [_F[op], int(f[1])]
        return 1
    if op == "ae":   # ('ae', 'le' | 'ge' | 'eq' | 'ne', var, k): AbstractElement over the store (formula.hpp:14-77)
        out += [_F[op], _AE[f[1]], int(f[2]), int(f[3])]
        return 1
    if op in ("true", "false"):   # ('true',) / ('false',): the constant formulas (formula.hpp:169-239)
        out.append(_F[op])
        return 1
    out.append(_F[op])
    if op in ("le", "gt", "eq", "ne"):
        if max(_encode_term(f[1], out), _encode_term(f[2], out)) > TREE_TERM_DEPTH:
            raise Unsupported(f"term deeper than {TREE_TERM_DEPTH} levels")
        return 1
    return 1 + max(_encode_formula(f[1], out), _encode_formula(f[2], out))


def encode_tree(f):
    """A formula as an LPC_PC_TREE propagator: (kind, [(w0, w1), ...] word pairs, 0, -1)."""
    words = []
    if _encode_formula(f, words) > TREE_FORM_DEPTH:
        raise Unsupported(f"connectives nested deeper than {TREE_FORM_DEPTH} levels")
    if len(words) % 2:
        words.append(0)
    return PC_TREE, list(zip(words[0::2], words[1::2])), 0, -1


def flatten(formulas, tree=True, bitset=False):
    """List of formula trees -> (props [n,5] int32, terms [m,2] int32). Shapes without a flat kind become LPC_PC_TREE
    propagators when `tree` is set, else `Unsupported` is raised. bitset: the table is for NBitset stores (Unary<Neg> and
    Binary<Sub> are not folded into coefficients, see _linear_terms)."""
    props, terms = [], []
    for f in formulas:
        try:
            kind, ts, rhs, bvar = flatten_one(f, bitset)
        except Unsupported:
            if not tree:
                raise
            kind, ts, rhs, bvar = encode_tree(f)
        props.append((kind, len(terms), len(ts), rhs, bvar))
        terms += ts
    return (np.asarray(props, dtype=np.int32).reshape(-1, 5), np.asarray(terms, dtype=np.int32).reshape(-1, 2))


def to_tree(kind, ts, rhs, bvar):
    """The formula tree a flat propagator stands for (inverse of flatten_one; used by generators and tests)."""
    def term(c, v):
        return ("var", v) if c == 1 else ("mul", ("const", c), ("var", v))
    if kind in (PC_LIN_LE, PC_REIF_LIN_LE, PC_LIN_GE, PC_LIN_GT, PC_LIN_EQ, PC_LIN_EQ_VAR):
        lhs = (term(*ts[0]) if len(ts) == 1 else ("add", term(*ts[0]), term(*ts[1])) if len(ts) == 2
               else ("sum",) + tuple(term(c, v) for c, v in ts))
        if kind == PC_LIN_GE:
            return ("le", ("const", rhs), lhs)
        if kind == PC_LIN_GT:
            return ("gt", lhs, ("const", rhs))
        if kind == PC_LIN_EQ:
            return ("eq", lhs, ("const", rhs))
        if kind == PC_LIN_EQ_VAR:
            if len(ts) == 1 and ts[0][0] == 1:
                raise Unsupported("x = z is the EQ kind")
            return ("eq", lhs, ("var", bvar))
        le = ("le", lhs, ("const", rhs))
        return le if kind == PC_LIN_LE else ("equiv", ("lit", bvar), le)
    if kind == PC_EQ:
        return ("eq", ("var", ts[0][1]), ("var", ts[1][1]))
    if kind == PC_ABS_EQ:
        return ("eq", ("abs", ("var", ts[0][1])), ("var", ts[1][1]))
    if kind == PC_NEQ:
        return ("ne", ("var", ts[0][1]), ("var", ts[1][1]) if len(ts) == 2 else ("const", rhs))
    if kind == PC_CLAUSE:
        lits = [("lit", v) if c > 0 else ("nlit", v) for c, v in ts]
        f = lits[-1]
        for l in reversed(lits[:-1]):
            f = ("or", l, f)
        return f
    raise Unsupported(kind)
