"""ctypes binding of the CPU checker (oracle/pir_oracle.cpp, oracle/pc_oracle.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. The product package never imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liblpc_oracle.so")


def build(force=False):
    """Compile the checker with the committed Makefile (g++ only, no dependencies)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".cpp")]
    stale = force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liblpc_oracle.so"], check=True, capture_output=True)
    return _LIB


class Stats(ctypes.Structure):
    _fields_ = [("has_changed", ctypes.c_int32), ("is_bot", ctypes.c_int32), ("sweeps", ctypes.c_int64),
                ("deductions", ctypes.c_int64), ("seconds", ctypes.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB)
        i32p = ctypes.POINTER(ctypes.c_int32)
        i64p = ctypes.POINTER(ctypes.c_int64)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        L.lpco_pir_deduce.argtypes = [i32p, ctypes.c_int32, i32p, i32p]
        L.lpco_pir_deduce.restype = ctypes.c_int
        L.lpco_pir_ask.argtypes = [i32p, ctypes.c_int32, i32p]
        L.lpco_pir_ask.restype = ctypes.c_int
        L.lpco_pir_clamp_reified.argtypes = [i32p, ctypes.c_int32, i32p, ctypes.c_int64, i32p]
        L.lpco_pir_clamp_reified.restype = None
        L.lpco_pir_fixpoint.argtypes = [i32p, ctypes.c_int32, i32p, ctypes.c_int64, ctypes.c_int32,
                                        ctypes.c_int64, ctypes.POINTER(Stats)]
        L.lpco_pir_fixpoint.restype = None
        L.lpco_pir_fixpoint_perm.argtypes = [i32p, ctypes.c_int32, i32p, ctypes.c_int64, i64p, ctypes.c_int32,
                                             ctypes.c_int64, ctypes.POINTER(Stats)]
        L.lpco_pir_fixpoint_perm.restype = None
        L.lpco_pir_ask_all.argtypes = [i32p, ctypes.c_int32, i32p, ctypes.c_int64, u8p]
        L.lpco_pir_ask_all.restype = ctypes.c_int64
        L.lpco_pir_batch_fixpoint.argtypes = [i32p, ctypes.c_int32, ctypes.c_int32, i32p, ctypes.c_int64,
                                              ctypes.c_int32, u8p, i32p, i64p]
        L.lpco_pir_batch_fixpoint.restype = ctypes.c_double
        L.lpco_pir_exhaustive_count.argtypes = [ctypes.c_int, ctypes.c_int]
        L.lpco_pir_exhaustive_count.restype = ctypes.c_int64
        L.lpco_pir_exhaustive.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          i64p, i32p]
        L.lpco_pir_exhaustive.restype = None
        L.lpco_div.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]
        L.lpco_div.restype = ctypes.c_int32
        _lib = L
    return _lib


def _p32(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def _store(store):
    s = np.ascontiguousarray(store, dtype=np.int32)
    assert s.ndim == 2 and s.shape[1] == 2, "store must be [nvars, 2] int32 {lb, ub}"
    return s


def _recs(records):
    r = np.ascontiguousarray(records, dtype=np.int32)
    assert r.ndim == 2 and r.shape[1] == 4, "records must be [n, 4] int32 {op, x, y, z}"
    return r


def pir_deduce(store, rec, is_bot=False):
    """One PIR::deduce(bytecode) step (pir.hpp:721-817). Returns (store', changed, is_bot)."""
    s = _store(store).copy()
    r = np.asarray(rec, dtype=np.int32).copy()
    b = ctypes.c_int32(int(is_bot))
    c = lib().lpco_pir_deduce(_p32(s), s.shape[0], _p32(r), ctypes.byref(b))
    return s, bool(c), bool(b.value)


def pir_ask(store, rec):
    s = _store(store)
    r = np.asarray(rec, dtype=np.int32).copy()
    return bool(lib().lpco_pir_ask(_p32(s), s.shape[0], _p32(r)))


def pir_clamp_reified(store, records):
    s = _store(store).copy()
    r = _recs(records)
    b = ctypes.c_int32(0)
    lib().lpco_pir_clamp_reified(_p32(s), s.shape[0], _p32(r), r.shape[0], ctypes.byref(b))
    return s


def pir_fixpoint(store, records, stop_on_bot=True, max_sweeps=0, perm=None):
    """Gauss-Seidel fixpoint (or a fixed-permutation chaotic iteration). Returns (store', Stats)."""
    s = _store(store).copy()
    r = _recs(records)
    st = Stats()
    if perm is None:
        lib().lpco_pir_fixpoint(_p32(s), s.shape[0], _p32(r), r.shape[0], int(stop_on_bot), max_sweeps,
                                ctypes.byref(st))
    else:
        p = np.ascontiguousarray(perm, dtype=np.int64)
        lib().lpco_pir_fixpoint_perm(_p32(s), s.shape[0], _p32(r), r.shape[0],
                                     p.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), int(stop_on_bot),
                                     max_sweeps, ctypes.byref(st))
    return s, st


def pir_ask_all(store, records, want_bits=False):
    s = _store(store)
    r = _recs(records)
    bits = np.zeros(r.shape[0], dtype=np.uint8) if want_bits else None
    n = lib().lpco_pir_ask_all(_p32(s), s.shape[0], _p32(r), r.shape[0],
                               bits.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)) if want_bits else None)
    return (int(n), bits) if want_bits else int(n)


def pir_batch_fixpoint(stores, records, threads=1):
    """stores [n_stores, nvars, 2]. Returns (stores', flags u8, sweeps i32, deductions, seconds)."""
    s = np.ascontiguousarray(stores, dtype=np.int32).copy()
    assert s.ndim == 3 and s.shape[2] == 2
    r = _recs(records)
    flags = np.zeros(s.shape[0], dtype=np.uint8)
    sweeps = np.zeros(s.shape[0], dtype=np.int32)
    ded = ctypes.c_int64(0)
    sec = lib().lpco_pir_batch_fixpoint(_p32(s), s.shape[0], s.shape[1], _p32(r), r.shape[0], threads,
                                        flags.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), _p32(sweeps),
                                        ctypes.byref(ded))
    return s, flags, sweeps, int(ded.value), float(sec)


SEARCH_COLS = ("solutions", "nodes", "fails", "best", "incomplete", "unknown_leaves")


def pir_search(roots, records, branch_vars, objective_var=-1, max_nodes=0, max_depth=64, threads=1):
    """Depth-first search (input order, bisection) around the Gauss-Seidel fixpoint, one tree per root store
    [n_stores, nvars, 2]. Returns int64 [n_stores, 6] with the columns SEARCH_COLS."""
    s = np.ascontiguousarray(roots, dtype=np.int32)
    assert s.ndim == 3 and s.shape[2] == 2
    r = _recs(records)
    bv = np.ascontiguousarray(branch_vars, dtype=np.int32)
    out = np.zeros((s.shape[0], 6), dtype=np.int64)
    f = lib().lpco_pir_search
    f.restype = None
    f.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                  ctypes.c_int32, ctypes.c_int32, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]
    f(s.ctypes.data, s.shape[0], s.shape[1], r.ctypes.data, r.shape[0], bv.ctypes.data, bv.shape[0], objective_var,
      max_nodes, max_depth, threads, out.ctypes.data)
    return out


def div(a, op, b):
    return int(lib().lpco_div(a, op, b))


EXH_KEYS = ("cases", "bot_cases", "unsound", "incomplete", "spurious_bot", "bad_entail", "entailed_cases",
            "not_converged")


def pir_exhaustive(op, minval, maxval, complete, threads=0, want_fixpoints=False):
    """The exhaustive bounds-consistency property (bound_consistency_test.hpp:155-225) on the oracle.
    Returns (dict of counters, fixpoints [cases, 7] or None)."""
    threads = threads or (os.cpu_count() or 1)
    n = lib().lpco_pir_exhaustive_count(minval, maxval)
    out = np.zeros(8, dtype=np.int64)
    fix = np.zeros((n, 7), dtype=np.int32) if want_fixpoints else None
    lib().lpco_pir_exhaustive(op, minval, maxval, int(complete), threads,
                              out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                              _p32(fix) if want_fixpoints else None)
    return dict(zip(EXH_KEYS, (int(v) for v in out))), fix


# ---- PC (oracle/pc_oracle.cpp) -------------------------------------------------------------------------------------
# prefix token codes of the formula stream (enum Tok in pc_oracle.cpp)
T_CONST, T_VAR, T_NEG, T_ABS, T_ADD, T_SUB, T_MUL, T_NARY_ADD, T_MIN, T_MAX = 1, 2, 3, 4, 5, 6, 7, 8, 9, 10
T_TDIV, T_FDIV, T_CDIV, T_EDIV, T_NARY_MUL, F_AE = 11, 12, 13, 14, 15, 31
AE_OPS = {"le": 0, "ge": 1, "eq": 2, "ne": 3}
F_VARLIT, F_NVARLIT, F_LEQ, F_GT, F_EQ, F_NEQ, F_AND, F_OR, F_EQUIV, F_IMPLY, F_XOR = 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30

_pc_bound = False


def _pc_lib():
    global _pc_bound
    L = lib()
    if not _pc_bound:
        i32p = ctypes.POINTER(ctypes.c_int32)
        L.lpco_pc_parse.argtypes = [i32p, ctypes.c_int32]
        L.lpco_pc_parse.restype = ctypes.c_void_p
        L.lpco_pc_free.argtypes = [ctypes.c_void_p]
        L.lpco_pc_free.restype = None
        L.lpco_pc_deduce.argtypes = [ctypes.c_void_p, ctypes.c_int32, i32p, ctypes.c_int32, i32p]
        L.lpco_pc_deduce.restype = ctypes.c_int
        L.lpco_pc_ask.argtypes = [ctypes.c_void_p, ctypes.c_int32, i32p, ctypes.c_int32]
        L.lpco_pc_ask.restype = ctypes.c_int
        L.lpco_pc_term_project.argtypes = [i32p, i32p, ctypes.c_int32, i32p]
        L.lpco_pc_term_project.restype = None
        L.lpco_pc_term_embed.argtypes = [i32p, i32p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]
        L.lpco_pc_term_embed.restype = ctypes.c_int
        L.lpco_pc_fixpoint.argtypes = [ctypes.c_void_p, i32p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int64,
                                       ctypes.POINTER(Stats)]
        L.lpco_pc_fixpoint.restype = None
        L.lpco_pc_ask_all.argtypes = [ctypes.c_void_p, i32p, ctypes.c_int32, ctypes.POINTER(ctypes.c_uint8)]
        L.lpco_pc_ask_all.restype = ctypes.c_int64
        u64p = ctypes.POINTER(ctypes.c_uint64)
        L.lpco_pc_deduce_bits.argtypes = [ctypes.c_void_p, ctypes.c_int32, u64p, ctypes.c_int32, i32p]
        L.lpco_pc_deduce_bits.restype = ctypes.c_int
        L.lpco_pc_ask_bits.argtypes = [ctypes.c_void_p, ctypes.c_int32, u64p, ctypes.c_int32]
        L.lpco_pc_ask_bits.restype = ctypes.c_int
        L.lpco_pc_fixpoint_bits.argtypes = [ctypes.c_void_p, u64p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int64,
                                            ctypes.POINTER(Stats)]
        L.lpco_pc_fixpoint_bits.restype = None
        L.lpco_pc_ask_all_bits.argtypes = [ctypes.c_void_p, u64p, ctypes.c_int32, ctypes.POINTER(ctypes.c_uint8)]
        L.lpco_pc_ask_all_bits.restype = ctypes.c_int64
        L.lpco_nbit.argtypes = [ctypes.c_int32, ctypes.c_int32]
        L.lpco_nbit.restype = ctypes.c_uint64
        L.lpco_nbit_bounds.argtypes = [ctypes.c_uint64, i32p]
        L.lpco_nbit_bounds.restype = None
        _pc_bound = True
    return L


def _p64(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))


def nbit(lb, ub):
    """NBitset<64>(lb, ub) (pc_bitset_test.cpp: NBit(1, 9))."""
    return int(_pc_lib().lpco_nbit(int(lb), int(ub)))


def nbit_from_set(values):
    """NBitset<64>::from_set (pc_bitset_test.cpp:81)."""
    b = 0
    for v in values:
        b |= nbit(v, v)
    return b


def nbit_bounds(bits):
    out = np.zeros(2, dtype=np.int32)
    _pc_lib().lpco_nbit_bounds(int(bits), _p32(out))
    return int(out[0]), int(out[1])


def nbit_store(store):
    """Interval store [n,2] -> uint64 cells."""
    s = _store(store)
    return np.array([nbit(l, u) for l, u in s.tolist()], dtype=np.uint64)


def flatten(tree):
    """Nested tuples -> prefix int list. ('var', v), ('const', k), ('neg', t), ('abs', t), ('add'|'sub'|'mul', a, b),
    ('sum', t1, ..., tn); formulas ('lit', v), ('nlit', v), ('le'|'gt'|'eq'|'ne', a, b), ('and'|'or'|'equiv', f, g)."""
    op = tree[0]
    unary = {"neg": T_NEG, "abs": T_ABS}
    binary = {"add": T_ADD, "sub": T_SUB, "mul": T_MUL, "min": T_MIN, "max": T_MAX, "tdiv": T_TDIV, "fdiv": T_FDIV,
              "cdiv": T_CDIV, "ediv": T_EDIV, "le": F_LEQ, "gt": F_GT, "eq": F_EQ,
              "ne": F_NEQ, "and": F_AND, "or": F_OR, "equiv": F_EQUIV, "imply": F_IMPLY, "xor": F_XOR}
    if op == "var":
        return [T_VAR, int(tree[1])]
    if op == "const":
        return [T_CONST, int(tree[1])]
    if op == "lit":
        return [F_VARLIT, int(tree[1])]
    if op == "nlit":
        return [F_NVARLIT, int(tree[1])]
    if op in unary:
        return [unary[op]] + flatten(tree[1])
    if op in binary:
        return [binary[op]] + flatten(tree[1]) + flatten(tree[2])
    if op == "ae":   # ('ae', 'le' | 'ge' | 'eq' | 'ne', var, k): a store-level element (AbstractElement)
        return [F_AE, AE_OPS[tree[1]], int(tree[2]), int(tree[3])]
    if op in ("true", "false"):   # the constant formulas (formula.hpp:169-239)
        return [32 if op == "true" else 33]
    if op in ("sum", "prod"):
        out = [T_NARY_ADD if op == "sum" else T_NARY_MUL, len(tree) - 1]
        for t in tree[1:]:
            out += flatten(t)
        return out
    raise ValueError(f"unknown node {op}")


class PCModel:
    """A list of PC propagators (formula trees) held by the CPU checker."""

    def __init__(self, formulas):
        self.formulas = list(formulas)
        stream = []
        for f in self.formulas:
            stream += flatten(f)
        self._stream = np.asarray(stream if stream else [0], dtype=np.int32)
        self._h = _pc_lib().lpco_pc_parse(_p32(self._stream), len(self.formulas))

    def __len__(self):
        return len(self.formulas)

    def deduce(self, i, store, is_bot=False):
        s = _store(store).copy()
        b = ctypes.c_int32(int(is_bot))
        c = _pc_lib().lpco_pc_deduce(self._h, i, _p32(s), s.shape[0], ctypes.byref(b))
        return s, bool(c), bool(b.value)

    def ask(self, i, store):
        s = _store(store)
        return bool(_pc_lib().lpco_pc_ask(self._h, i, _p32(s), s.shape[0]))

    def fixpoint(self, store, stop_on_bot=True, max_sweeps=0):
        s = _store(store).copy()
        st = Stats()
        _pc_lib().lpco_pc_fixpoint(self._h, _p32(s), s.shape[0], int(stop_on_bot), max_sweeps, ctypes.byref(st))
        return s, st

    def ask_all(self, store, want_bits=False):
        s = _store(store)
        bits = np.zeros(max(1, len(self)), dtype=np.uint8)
        n = _pc_lib().lpco_pc_ask_all(self._h, _p32(s), s.shape[0], bits.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
        return (int(n), bits[:len(self)]) if want_bits else int(n)

    # the same model over a VStore<NBitset<64>>: cells = uint64 [nvars]
    def deduce_bits(self, i, cells, is_bot=False):
        s = np.ascontiguousarray(cells, dtype=np.uint64).copy()
        b = ctypes.c_int32(int(is_bot))
        c = _pc_lib().lpco_pc_deduce_bits(self._h, i, _p64(s), s.shape[0], ctypes.byref(b))
        return s, bool(c), bool(b.value)

    def fixpoint_bits(self, cells, stop_on_bot=True, max_sweeps=0):
        s = np.ascontiguousarray(cells, dtype=np.uint64).copy()
        st = Stats()
        _pc_lib().lpco_pc_fixpoint_bits(self._h, _p64(s), s.shape[0], int(stop_on_bot), max_sweeps, ctypes.byref(st))
        return s, st

    def ask_all_bits(self, cells, want_bits=False):
        s = np.ascontiguousarray(cells, dtype=np.uint64)
        bits = np.zeros(max(1, len(self)), dtype=np.uint8)
        n = _pc_lib().lpco_pc_ask_all_bits(self._h, _p64(s), s.shape[0], bits.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
        return (int(n), bits[:len(self)]) if want_bits else int(n)

    def __del__(self):
        if getattr(self, "_h", None):
            _pc_lib().lpco_pc_free(self._h)
            self._h = None


def pc_term_project(term, store):
    s = _store(store)
    stream = np.asarray(flatten(term), dtype=np.int32)
    out = np.zeros(2, dtype=np.int32)
    _pc_lib().lpco_pc_term_project(_p32(stream), _p32(s), s.shape[0], _p32(out))
    return int(out[0]), int(out[1])


def pc_term_embed(term, store, lb, ub):
    s = _store(store).copy()
    stream = np.asarray(flatten(term), dtype=np.int32)
    c = _pc_lib().lpco_pc_term_embed(_p32(stream), _p32(s), s.shape[0], lb, ub)
    return s, bool(c)
