"""Synthetic propagator networks of the shapes named in BASELINE.json.

Pure numpy, deterministic (counter-based splitmix64; seed = 0xB2000000 + config id). Every network has a
*planted solution*: a hidden value s[v] per variable, only records that hold under s are emitted, and every domain
contains s[v], so propagation happens but the network can never fail. `failing_twin` moves one domain off its
planted value so that the bot flag can be checked too. Magnitudes are bounded (|s| <= 4096, MUL operands
|s| <= 64) so that no int32 intermediate overflows (the reference has UB there, pir.hpp:759-772).

Variable classes (kept apart so that the network does not solve itself in one sweep — see `pir_network`):
  A  addition population: x, y, z of `x = y + z`
  M  small multiplication operands, P  products: `x = y * z` with y, z in M and x in P
  B  0/1 variables: results of the reified `b = (y <= z)` / `b = (y == z)` over A, M and P

Records are returned sorted by (op, y, x, z), the order PIR::deduce(tell) establishes (pir.hpp:343-347).
"""
import numpy as np

ADD, MUL, MIN, MAX, TDIV, FDIV, CDIV, EDIV, EQ, LEQ = 2, 4, 6, 7, 25, 27, 29, 31, 46, 48
SEED_BASE = 0xB2000000
CLS_A, CLS_M, CLS_P, CLS_B = 0, 1, 2, 3


def splitmix64(seed, stream, n):
    """n uint64 values of stream `stream`: counter-based, so any slice is reproducible independently."""
    with np.errstate(over="ignore"):
        x = (np.uint64(seed) + np.uint64(stream) * np.uint64(0xD1342543DE82EF95)
             + (np.arange(1, n + 1, dtype=np.uint64)) * np.uint64(0x9E3779B97F4A7C15))
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return x


class _Rng:
    def __init__(self, seed):
        self.seed, self.stream = seed, 0

    def u64(self, n):
        self.stream += 1
        return splitmix64(self.seed, self.stream, n)

    def below(self, n, bound):
        """n integers in [0, bound) (bound may be an array)."""
        return (self.u64(n) % np.asarray(bound, dtype=np.uint64)).astype(np.int64)

    def between(self, n, lo, hi):
        """n integers in [lo, hi] inclusive (arrays allowed)."""
        lo = np.asarray(lo, dtype=np.int64)
        hi = np.asarray(hi, dtype=np.int64)
        return lo + self.below(n, hi - lo + 1)


def sort_records(recs):
    order = np.lexsort((recs[:, 3], recs[:, 1], recs[:, 2], recs[:, 0]))
    return np.ascontiguousarray(recs[order])


class Network:
    """records [n,4] int32 {op,x,y,z}; store [nvars,2] int32 {lb,ub}; solution [nvars] planted values."""

    def __init__(self, records, store, solution, meta):
        self.records, self.store, self.solution, self.meta = records, store, solution, meta
        self.nvars = store.shape[0]

    def failing_twin(self):
        """Same network with the result variable of the first ADD record pinned one off the value that its
        (pinned) operands force: the fixpoint must end at bot."""
        st = self.store.copy()
        i = int(np.flatnonzero(self.records[:, 0] == ADD)[0])
        _, x, y, z = self.records[i]
        for v in (y, z):
            st[v] = (self.solution[v], self.solution[v])
        st[x] = (self.solution[x] + 1, self.solution[x] + 1)
        return Network(self.records, st, self.solution, dict(self.meta, failing=True))


class _ValueIndex:
    """Variables of one class indexed by planted value: lookup of a variable with a given value near a position."""

    def __init__(self, members, s, nvars, base):
        self.nvars, self.base = nvars, base
        key = (s[members] + base) * nvars + members
        o = np.argsort(key, kind="stable")
        self.keys, self.idx = key[o], members[o]

    def find(self, value, pos):
        want = value + self.base
        q = want * self.nvars + pos
        j = np.clip(np.searchsorted(self.keys, q), 0, len(self.keys) - 1)
        jm = np.clip(j - 1, 0, len(self.keys) - 1)
        ok_j = self.keys[j] // self.nvars == want
        ok_m = self.keys[jm] // self.nvars == want
        dj = np.abs(self.keys[j] % self.nvars - pos)
        dm = np.abs(self.keys[jm] % self.nvars - pos)
        use_m = ok_m & (~ok_j | (dm < dj))
        out = np.where(use_m, self.idx[jm], self.idx[j])
        return np.where(ok_j | ok_m, out, -1)


def pir_network(nvars, nrec, seed, mix=((ADD, 0.5), (MUL, 0.25), (LEQ, 0.25)), window=4096, local_frac=0.8, width=16,
                singleton_frac=0.01, class_frac=(0.60, 0.20, 0.15, 0.05), bool_fixed_frac=0.30, value_range=None):
    """Planted-solution PIR network (configs 1, 2, 4 of BASELINE.json).

    window/local_frac: the locality knob — that fraction of a record's other operands lies within +-window of y.

    Why classes and slack >= 1: `x = y + z` copies tightness (slack(x) <= slack(y) + slack(z)); with ~15 incidences
    per variable a zero-slack fraction above ~1/60 percolates and a single Gauss-Seidel sweep solves the whole
    network (measured: 10 % singletons -> 99.9 % singletons after one sweep). Multiplication pins its small operands
    almost immediately, so it gets its own operand / product populations; what remains is an addition network
    whose bounds tighten over ~10 sweeps, coupled to the rest through the reified comparisons.
    """
    rng = _Rng(seed)
    R = int(value_range or min(4096, max(16, nvars // 64)))        # |s| <= R
    r_small = int(np.floor(np.sqrt(R)))             # MUL operands
    cum = np.cumsum(np.asarray(class_frac, dtype=np.float64))
    u = rng.below(nvars, 1 << 20) / float(1 << 20)
    cls = np.searchsorted(cum / cum[-1], u, side="right").clip(0, 3)
    allv = np.arange(nvars, dtype=np.int64)
    members = [allv[cls == c] for c in range(4)]
    for c in range(4):
        assert len(members[c]) >= 4, "network too small for four variable classes"
    s = rng.between(nvars, -R, R)
    s = np.where(cls == CLS_M, rng.between(nvars, -r_small, r_small), s)
    s = np.where(cls == CLS_P, rng.between(nvars, -r_small, r_small) * rng.between(nvars, -r_small, r_small), s)
    s = np.where(cls == CLS_B, rng.below(nvars, 2), s)
    idx_a = _ValueIndex(members[CLS_A], s, nvars, R)
    idx_p = _ValueIndex(members[CLS_P], s, nvars, R)
    idx_b = _ValueIndex(members[CLS_B], s, nvars, 0)
    ints = allv[cls != CLS_B]

    def near(pos, n):
        """n positions near `pos`: within +-window with probability local_frac, anywhere otherwise."""
        loc = rng.below(n, 1000) < int(local_frac * 1000)
        off = rng.between(n, -window, window)
        return np.where(loc, np.clip(pos + off, 0, nvars - 1), rng.below(n, nvars))

    def member_at(m, pos):
        """The member of the sorted index list `m` at or after position `pos`."""
        return m[np.clip(np.searchsorted(m, pos), 0, len(m) - 1)]

    parts = []
    for op, frac in mix:
        want = int(round(nrec * frac))
        got, have = [], 0
        while have < want:
            n = int((want - have) * 1.6) + 64
            if op == ADD:        # x = y + z over class A: pick y and x, look z up by value
                y = member_at(members[CLS_A], rng.below(n, nvars))
                x = member_at(members[CLS_A], near(y, n))
                d = s[x] - s[y]
                z = idx_a.find(d, near(y, n))
                ok = (np.abs(d) <= R) & (z >= 0)
            elif op == MUL:      # x = y * z: operands from class M, product looked up in class P
                y = member_at(members[CLS_M], rng.below(n, nvars))
                z = member_at(members[CLS_M], near(y, n))
                p = s[y] * s[z]
                x = idx_p.find(p, near(y, n))
                ok = (np.abs(p) <= R) & (x >= 0)
            elif op in (LEQ, EQ):    # b = (y <= z) / b = (y == z), b from the 0/1 pool
                y = member_at(ints, rng.below(n, nvars))
                z = member_at(ints, near(y, n))
                truth = (s[y] <= s[z]) if op == LEQ else (s[y] == s[z])
                x = idx_b.find(truth.astype(np.int64), near(y, n))
                ok = x >= 0
            elif op in (MIN, MAX):
                y = member_at(members[CLS_A], rng.below(n, nvars))
                z = member_at(members[CLS_A], near(y, n))
                val = np.minimum(s[y], s[z]) if op == MIN else np.maximum(s[y], s[z])
                x = idx_a.find(val, near(y, n))
                ok = x >= 0
            else:
                raise ValueError(f"op {op} not supported by this generator")
            ok &= (x != y) & (x != z) & (y != z)
            rec = np.stack([np.full(n, op, dtype=np.int64), x, y, z], axis=1)[ok]
            got.append(rec[: want - have])
            have += len(got[-1])
        parts.append(np.concatenate(got))
    recs = np.concatenate(parts).astype(np.int32)
    a = rng.between(nvars, 1, width)
    b = rng.between(nvars, 1, width)
    single = rng.below(nvars, 10000) < int(singleton_frac * 10000)
    lb = np.where(single, s, s - a)
    ub = np.where(single, s, s + b)
    bfix = rng.below(nvars, 1000) < int(bool_fixed_frac * 1000)
    lb = np.where(cls == CLS_B, np.where(bfix, s, 0), lb)
    ub = np.where(cls == CLS_B, np.where(bfix, s, 1), ub)
    store = np.stack([lb, ub], axis=1).astype(np.int32)
    recs = sort_records(recs)
    assert check_solution(recs, s)
    return Network(recs, store, s.astype(np.int32), dict(nvars=nvars, nrec=len(recs), seed=seed, R=R))


def check_solution(recs, s):
    """Every record holds under the planted assignment."""
    op, x, y, z = (recs[:, i].astype(np.int64) for i in range(4))
    sx, sy, sz = s[x], s[y], s[z]
    ok = np.ones(len(recs), dtype=bool)
    ok &= np.where(op == ADD, sx == sy + sz, True)
    ok &= np.where(op == MUL, sx == sy * sz, True)
    ok &= np.where(op == LEQ, sx == (sy <= sz), True)
    ok &= np.where(op == EQ, sx == (sy == sz), True)
    ok &= np.where(op == MIN, sx == np.minimum(sy, sz), True)
    ok &= np.where(op == MAX, sx == np.maximum(sy, sz), True)
    return bool(ok.all())


def config1():
    """10k vars, 50k ternary propagators (x=y+z, x=y*z, x=(y<=z)) — the reference's CPU-runnable case."""
    return pir_network(10_000, 50_000, SEED_BASE + 1)


def config2(scale=1.0):
    """1M vars, 5M ternary propagators; `scale` shrinks it for parity tests."""
    return pir_network(int(1_000_000 * scale), int(5_000_000 * scale), SEED_BASE + 2)


def config4_base():
    """Base model of the batched EPS config: 2k vars / 10k propagators + 16 decision variables + objective.
    Decision variables are the widest addition-class variables of the propagated base model."""
    return pir_network(2_000, 10_000, SEED_BASE + 4, window=256, value_range=1024)


def eps_decisions(records, root_store, n=16, min_degree=8):
    """n decision variables — the widest domains of the root fixpoint among variables with at least `min_degree`
    incident propagators (ties by index) — and an objective variable (the next one in that order)."""
    nvars = len(root_store)
    w = root_store[:, 1].astype(np.int64) - root_store[:, 0]
    deg = np.bincount(np.asarray(records)[:, 1:].ravel(), minlength=nvars)
    cand = np.flatnonzero(deg >= min_degree)
    order = cand[np.lexsort((cand, -w[cand]))]
    assert len(order) > n
    return sorted(int(v) for v in order[:n]), int(order[n])


class PcNetwork:
    """props [n,5] int32 {kind, first_term, n_terms, rhs, bvar}, terms [m,2] int32 {coef, var} (include/lpc_pc.h),
    store [nvars,2], solution [nvars]."""

    def __init__(self, props, terms, store, solution, meta):
        self.props, self.terms, self.store, self.solution, self.meta = props, terms, store, solution, meta
        self.nvars = store.shape[0]

    def formulas(self):
        """The formula trees the flat propagators stand for (for the tree-walking checker)."""
        from . import pcflat
        out = []
        for kind, first, n, rhs, bvar in self.props.tolist():
            ts = [tuple(t) for t in self.terms[first:first + n].tolist()]
            out.append(pcflat.to_tree(kind, ts, rhs, bvar))
        return out


def config3(scale=1.0, seed=SEED_BASE + 3, reified_frac=0.30):
    """PC with n-ary linear sums and reified sums (BASELINE.json configs[2]): 200k vars, constraints
    sum c_i x_i <= k and b <=> (sum c_i x_i <= k), arity U{2..8}, c_i in {1,2,3,5}, until 1M terms are emitted."""
    nvars = int(200_000 * scale)
    want_terms = int(1_000_000 * scale)
    rng = _Rng(seed)
    nbool = max(16, nvars // 20)
    is_bool = np.zeros(nvars, dtype=bool)
    is_bool[rng.below(nbool, nvars)] = True
    ints = np.flatnonzero(~is_bool)
    bools = np.flatnonzero(is_bool)
    s = np.where(is_bool, rng.below(nvars, 2), rng.between(nvars, 0, 100))
    ncons = int(want_terms / 5.0 * 1.05) + 8
    arity = rng.between(ncons, 2, 8)
    cum = np.cumsum(arity)
    ncons = int(np.searchsorted(cum, want_terms)) + 1
    arity = arity[:ncons]
    first = np.concatenate([[0], np.cumsum(arity)[:-1]])
    m = int(arity.sum())
    # variables of a constraint: a random anchor, then distinct neighbours (anchor + a strictly increasing offset)
    anchor = np.repeat(rng.below(ncons, len(ints)), arity)
    pos_in = np.arange(m) - np.repeat(first, arity)
    step = rng.between(m, 1, 64)
    csum = np.cumsum(step)
    off = np.where(pos_in == 0, 0, csum - np.repeat(csum[first], arity))   # strictly increasing inside a constraint
    var = ints[(anchor + off) % len(ints)]
    coef = np.array([1, 2, 3, 5], dtype=np.int64)[rng.below(m, 4)]
    lhs = np.add.reduceat(coef * s[var], first)
    reified = rng.below(ncons, 1000) < int(reified_frac * 1000)
    # right-hand sides: non-reified constraints hold with a small slack; reified ones sit around the planted sum
    slack = rng.between(ncons, 0, 12)
    rhs = np.where(reified, lhs + rng.between(ncons, -10, 10), lhs + slack)
    truth = lhs <= rhs
    bvar = np.full(ncons, -1, dtype=np.int64)
    # reification variable with the right planted value, near the anchor
    idx_b = _ValueIndex(bools.astype(np.int64), s, nvars, 0)
    pick = idx_b.find(truth.astype(np.int64), var[first])
    bvar = np.where(reified, pick, -1)
    assert (bvar[reified] >= 0).all()
    kind = np.where(reified, 2, 1)
    props = np.stack([kind, first, arity, rhs, bvar], axis=1).astype(np.int32)
    terms = np.stack([coef, var], axis=1).astype(np.int32)
    a = rng.between(nvars, 1, 16)
    b = rng.between(nvars, 1, 16)
    lb = np.maximum(0, s - a)
    ub = s + b
    bfix = rng.below(nvars, 1000) < 300
    lb = np.where(is_bool, np.where(bfix, s, 0), lb)
    ub = np.where(is_bool, np.where(bfix, s, 1), ub)
    store = np.stack([lb, ub], axis=1).astype(np.int32)
    # sort by (kind, length) like PC::deduce(tell) (pc.hpp:636-643): stable, so equal keys keep emission order
    order = np.lexsort((props[:, 2], props[:, 0]))
    props = np.ascontiguousarray(props[order])
    return PcNetwork(props, terms, store, s.astype(np.int32), dict(nvars=nvars, nprops=len(props), nterms=m, seed=seed))


def config5(scale=1.0, seed=SEED_BASE + 5):
    """Bitset-domain PC shapes (tests/pc_bitset_test.cpp) scaled up (BASELINE.json configs[4]): 100k vars with domains
    inside [0, 61], 500k propagators: 40 % x != y, 30 % x = y, 25 % 4-literal clauses, 5 % y = |x|."""
    nvars = int(100_000 * scale)
    nprops = int(500_000 * scale)
    rng = _Rng(seed)
    nbool = max(16, nvars // 10)
    is_bool = np.zeros(nvars, dtype=bool)
    is_bool[rng.below(nbool, nvars)] = True
    ints = np.flatnonzero(~is_bool).astype(np.int64)
    bools = np.flatnonzero(is_bool).astype(np.int64)
    s = np.where(is_bool, rng.below(nvars, 2), rng.between(nvars, 0, 61))
    idx_i = _ValueIndex(ints, s, nvars, 0)
    n_ne, n_eq, n_cl = int(nprops * 0.40), int(nprops * 0.30), int(nprops * 0.25)
    n_abs = nprops - n_ne - n_eq - n_cl
    props, terms = [], []

    def pairs(n, same):
        out_x, out_y = [], []
        have = 0
        while have < n:
            k = int((n - have) * 1.3) + 64
            x = ints[rng.below(k, len(ints))]
            if same:
                y = idx_i.find(s[x], np.clip(x + rng.between(k, -2048, 2048), 0, nvars - 1))
                ok = (y >= 0) & (y != x)
            else:
                y = ints[np.clip(np.searchsorted(ints, np.clip(x + rng.between(k, -2048, 2048), 0, nvars - 1)), 0, len(ints) - 1)]
                ok = s[x] != s[y]
            out_x.append(x[ok][: n - have]); out_y.append(y[ok][: n - have])
            have += len(out_x[-1])
        return np.concatenate(out_x), np.concatenate(out_y)

    t0 = 0
    for kind, n, same in ((4, n_ne, False), (3, n_eq, True), (6, n_abs, True)):
        x, y = pairs(n, same)
        first = t0 + 2 * np.arange(n)
        props.append(np.stack([np.full(n, kind), first, np.full(n, 2), np.zeros(n, dtype=np.int64), np.full(n, -1)], axis=1))
        terms.append(np.stack([np.ones(2 * n, dtype=np.int64), np.stack([x, y], axis=1).ravel()], axis=1))
        t0 += 2 * n
    # clauses: 4 literals over 0/1 variables, at least one true under the planted assignment
    v = bools[rng.below(4 * n_cl, len(bools))].reshape(n_cl, 4)
    sign = np.where(rng.below(4 * n_cl, 2).reshape(n_cl, 4) == 1, 1, -1)
    true_lit = (sign > 0) == (s[v] == 1)
    none = ~true_lit.any(axis=1)
    sign[none, 0] = np.where(s[v[none, 0]] == 1, 1, -1)     # make the first literal true
    first = t0 + 4 * np.arange(n_cl)
    props.append(np.stack([np.full(n_cl, 5), first, np.full(n_cl, 4), np.zeros(n_cl, dtype=np.int64), np.full(n_cl, -1)], axis=1))
    terms.append(np.stack([sign.ravel(), v.ravel()], axis=1))
    props = np.concatenate(props).astype(np.int32)
    terms = np.concatenate(terms).astype(np.int32)
    a = rng.between(nvars, 1, 12)
    b = rng.between(nvars, 1, 12)
    single = rng.below(nvars, 1000) < 50
    lb = np.where(single, s, np.maximum(0, s - a))
    ub = np.where(single, s, np.minimum(61, s + b))
    bfix = rng.below(nvars, 1000) < 300
    lb = np.where(is_bool, np.where(bfix, s, 0), lb)
    ub = np.where(is_bool, np.where(bfix, s, 1), ub)
    store = np.stack([lb, ub], axis=1).astype(np.int32)
    order = np.lexsort((props[:, 2], props[:, 0]))
    props = np.ascontiguousarray(props[order])
    return PcNetwork(props, terms, store, s.astype(np.int32), dict(nvars=nvars, nprops=len(props), nterms=len(terms), seed=seed))


def eps_stores(base_store, decision_vars, first_id, n, ids=None):
    """Host restatement of lpc_batch_init_split(_ids) (include/lpc.h): subproblem id bit j halves variable d_j."""
    out = np.repeat(base_store[None, :, :], n, axis=0).copy()
    ids = first_id + np.arange(n, dtype=np.int64) if ids is None else np.asarray(ids, dtype=np.int64)
    assert ids.shape == (n,)
    for j, v in enumerate(decision_vars):
        lb, ub = int(base_store[v, 0]), int(base_store[v, 1])
        mid = lb + ((ub - lb) >> 1)
        bit = (ids >> j) & 1
        out[:, v, 0] = np.where(bit == 1, mid + 1, lb)
        out[:, v, 1] = np.where(bit == 1, ub, mid)
    return out
