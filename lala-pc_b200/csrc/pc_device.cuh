// pc_device.cuh — the flat PC propagators (include/lpc_pc.h) on registers, for sm_100a.
//
// Device counterpart of PC::deduce(int) -> pc::Formula::deduce -> pc::Term::{project, embed}
// (lala-pc include/lala/pc.hpp:671-680, formula.hpp, terms.hpp) for the tree shapes named in lpc_pc.h. Every routine
// cites the tree walk it replaces. Bound arithmetic follows lala-core's Interval::project (un-vendored; the same
// restatement the CPU checker of the test-suite uses): ADD / SUB componentwise on the bounds, an infinite operand bound gives an infinite result
// bound, MUL / EDIV by a constant as the hull of the corner results. `Acc` abstracts the store: `load(v)` returns the
// current domain, `embed(v, itv)` joins (VStore::embed) and returns bit0 = changed, bit1 = became empty.
#pragma once
#include "pir_device.cuh"

namespace lpc {

enum PcKind : int { PC_LIN_LE = 1, PC_REIF_LIN_LE = 2, PC_EQ = 3, PC_NEQ = 4, PC_CLAUSE = 5, PC_ABS_EQ = 6,
                    PC_LIN_GE = 7, PC_LIN_GT = 8, PC_LIN_EQ = 9, PC_LIN_EQ_VAR = 10, PC_TREE = 11 };
__host__ __device__ __forceinline__ bool pc_is_linear(int kind) { return kind == PC_LIN_LE || kind == PC_REIF_LIN_LE || (kind >= PC_LIN_GE && kind <= PC_LIN_EQ_VAR); }
// The general tree propagator (pc_tree.cuh); its stream lives in the term array. On the device the interpreter is its
// own translation unit (pc_tree.cu, built with a register cap so that the kernels can call it through the device ABI
// whatever their launch bounds are) and is entered through the global-memory accessor below.
struct UItv;   // the universes of the tree interpreter (pc_tree.cuh); an accessor names its own as Acc::Univ
struct UNb;
#ifdef LPC_HOST_HARNESS
template <class Acc> LPC_HD int pc_tree_deduce(Acc& a, const int* words);
template <class Acc> LPC_HD bool pc_tree_ask(const Acc& a, const int* words);
#else
// VStore<Interval<ZLB>> in global memory: gathers + lattice joins at L2.
struct GlobalAcc {
  typedef UItv Univ;
  int2* s;
  mutable int seen_bot;
  int touched;   // some embed tightened the store: what the fixpoint kernel votes with (a tree propagator's own return
                 // value can hide a change, formula.hpp:803)
  __device__ __forceinline__ Itv load(int v) const {
    const int2 d = __ldcg(&s[v]);
    seen_bot |= d.x > d.y;
    return Itv(d.x, d.y);
  }
  __device__ __forceinline__ int embed(int v, const Itv& u) {
    const int2 old = __ldcg(&s[v]);
    int f = 0;
    if(u.lb > old.x) { atomicMax(&s[v].x, u.lb); f = 1; }
    if(u.ub < old.y) { atomicMin(&s[v].y, u.ub); f = 1; }
    if(f && max(u.lb, old.x) > min(u.ub, old.y)) f |= 2;
    touched |= f;
    return f;
  }
};
__device__ int pc_tree_deduce_global(GlobalAcc& a, const int* words);
__device__ bool pc_tree_ask_global(const GlobalAcc& a, const int* words);
__device__ __forceinline__ int pc_tree_deduce(GlobalAcc& a, const int* words) { return pc_tree_deduce_global(a, words); }
__device__ __forceinline__ bool pc_tree_ask(const GlobalAcc& a, const int* words) { return pc_tree_ask_global(a, words); }
#endif
__host__ __device__ __forceinline__ bool pc_has_extra_lane(int kind) { return kind == PC_REIF_LIN_LE || kind == PC_LIN_EQ_VAR; }

struct PcTableDev {
  const int4* hdr;     // {kind | n_terms << 8, first_term, rhs, bvar}
  const int2* terms;   // {coef, var}
  long long n;
  long long n_terms;
  int nvars;
  // lane tiles of the fixpoint kernel (pc_fixpoint.cu): 32 x {coef, var, meta, rhs} per tile
  const int4* tiles;
  const int* tile_prop0;   // index of the first propagator of each tile
  long long n_tiles;
  const int* big;          // propagators with more than 32 lanes
  int n_big;
};

// lane meta word: bits 0-4 first lane of the propagator, 5-10 its lane count (1..32), 11-14 kind (0 = padding),
// bit 15 = this lane is the extra variable of the propagator (the Boolean of a reified sum, the z of sum = z)
#define PC_META(start, len, kind, isb) ((start) | ((len) << 5) | ((kind) << 11) | ((isb) << 15))
#define PC_META_KIND(m) (((m) >> 11) & 15)
#define PC_META_EXTRA(m) (((m) >> 15) & 1)

LPC_HD bool b_inf(int x) { return x == LPC_INF || x == LPC_MINF; }
LPC_HD int b_clamp(long long x) { return x >= LPC_INF ? LPC_INF : (x <= LPC_MINF ? LPC_MINF : (int)x); }
LPC_HD int b_add(int a, int b) { if(b_inf(a)) return a; if(b_inf(b)) return b; return b_clamp((long long)a + b); }
LPC_HD int b_neg(int a) { return a == LPC_INF ? LPC_MINF : (a == LPC_MINF ? LPC_INF : -a); }
LPC_HD int b_sub(int a, int b) { return b_add(a, b_neg(b)); }
LPC_HD int b_mul(int a, int b) {
  if(a == 0 || b == 0) return 0;
  if(b_inf(a) || b_inf(b)) return ((a < 0) != (b < 0)) ? LPC_MINF : LPC_INF;
  return b_clamp((long long)a * b);
}
LPC_HD int b_ediv(int a, int b) {   // Euclidean, b != 0 and finite
  if(b_inf(a)) return (b > 0) ? a : b_neg(a);
  long long q = (long long)a / b, r = (long long)a % b;
  if(r < 0) q += (b > 0) ? -1 : 1;
  return b_clamp(q);
}
LPC_HD bool contains0(const Itv& a) { return a.lb <= 0 && 0 <= a.ub; }            // a >= eq_zero
LPC_HD bool sub_of_zero(const Itv& a) { return a.is_bot() || (a.lb >= 0 && a.ub <= 0); }   // a <= eq_zero

// Binary<GroupMul>(Constant c, Variable x)::project / Variable::project (terms.hpp:399-405, 69-71)
LPC_HD Itv term_project(int c, const Itv& x) {
  if(c == 1) return x;
  if(x.is_bot()) return itv_bot();
  const int p = b_mul(c, x.lb), q = b_mul(c, x.ub);
  return Itv(min(p, q), max(p, q));
}
// ...::embed: the variable's residual for `c * x <= u` in the lattice sense (terms.hpp:384-396, 249-253)
LPC_HD Itv term_residual(int c, const Itv& u) {
  if(c == 1) return u;
  if(u.is_bot()) return itv_bot();
  const int p = b_ediv(u.lb, c), q = b_ediv(u.ub, c);
  return Itv(min(p, q), max(p, q));
}

// Nary<Add>::project (terms.hpp:465-478)
template <class Acc>
LPC_HD Itv lin_project(const Acc& a, const int2* terms, int n) {
  int2 t0 = terms[0];
  Itv accu = term_project(t0.x, a.load(t0.y));
  for(int i = 1; i < n; ++i) {
    const int2 t = terms[i];
    const Itv ti = term_project(t.x, a.load(t.y));
    accu = Itv(b_add(accu.lb, ti.lb), b_add(accu.ub, ti.ub));
  }
  return accu;
}

// Nary<Add>::embed(u) (terms.hpp:480-499): all = sum once; per term the others' sum through additive_inverse
// (GroupAdd::rev_op, :190-194), residual = u - others (left_residual, :196-198), then the term's own embed.
// The interpreter builds a lone term as itself, two terms as Binary<GroupAdd> and three or more as Nary<Add>
// (pc.hpp:277-296), and the three differ once infinite bounds are involved, so each arity keeps its own walk.
template <class Acc>
LPC_HD int lin_embed(Acc& a, const int2* terms, int n, const Itv& u, const Itv& all) {
  int f = 0;
  if(n == 1) return a.embed(terms[0].y, term_residual(terms[0].x, u));   // Variable / Binary<Mul>::embed
  if(n == 2) {   // Binary<GroupAdd>::embed (terms.hpp:376-397): x <- u - y, then y <- u - x with x re-read
    const int2 t0 = terms[0], t1 = terms[1];
    const Itv yt = term_project(t1.x, a.load(t1.y));
    f |= a.embed(t0.y, term_residual(t0.x, Itv(b_sub(u.lb, yt.ub), b_sub(u.ub, yt.lb))));
    const Itv xt = term_project(t0.x, a.load(t0.y));
    f |= a.embed(t1.y, term_residual(t1.x, Itv(b_sub(u.lb, xt.ub), b_sub(u.ub, xt.lb))));
    return f;
  }
  for(int i = 0; i < n; ++i) {
    const int2 t = terms[i];
    const Itv ti = term_project(t.x, a.load(t.y));
    const Itv others(b_add(all.lb, b_neg(ti.lb)), b_add(all.ub, b_neg(ti.ub)));
    const Itv res(b_sub(u.lb, others.ub), b_sub(u.ub, others.lb));
    f |= a.embed(t.y, term_residual(t.x, res));
  }
  return f;
}

// VariableLiteral (formula.hpp:97-120): ask / nask / deduce / contradeduce with `neg` folded in.
LPC_HD bool lit_ask(bool neg, const Itv& b) { return neg ? sub_of_zero(b) : !contains0(b); }
template <class Acc> LPC_HD int lit_deduce(Acc& a, bool neg, int v) { return a.embed(v, neg ? Itv(0, 0) : Itv(1, 1)); }

// PC::deduce(i) for one flat propagator. Returns bit0 = changed, bit1 = some variable became empty.
template <class Acc>
LPC_HD int pc_deduce(Acc& a, const int4 h, const int2* terms) {
  const int kind = h.x & 0xff, n = h.x >> 8, rhs = h.z, bvar = h.w;
  switch(kind) {
    case PC_LIN_LE: {   // Inequality<false>::deduce, right side constant (formula.hpp:796-800)
      const Itv all = lin_project(a, terms, n);
      return lin_embed(a, terms, n, Itv(LPC_MINF, rhs), all);
    }
    case PC_REIF_LIN_LE: {   // Biconditional::deduce (formula.hpp:421-427)
      const Itv b = a.load(bvar);
      const Itv all = lin_project(a, terms, n);
      if(lit_ask(false, b)) return lin_embed(a, terms, n, Itv(LPC_MINF, rhs), all);                 // g.deduce
      else if(lit_ask(true, b)) return lin_embed(a, terms, n, Itv(b_add(rhs, 1), LPC_INF), all);    // g.contradeduce: l > r (:779-785)
      else if(all.ub <= rhs) return lit_deduce(a, false, bvar);                                      // g.ask (:769)
      else if(all.lb > rhs) return lit_deduce(a, true, bvar);                                        // g.nask (:766)
      return 0;
    }
    case PC_LIN_GE: {   // Inequality<false>::deduce, left side constant (formula.hpp:801-804)
      const Itv all = lin_project(a, terms, n);
      return lin_embed(a, terms, n, Itv(rhs, LPC_INF), all);
    }
    case PC_LIN_GT: {   // Inequality<true>::deduce, right side constant (formula.hpp:779-785)
      const Itv all = lin_project(a, terms, n);
      return lin_embed(a, terms, n, Itv(b_add(rhs, 1), LPC_INF), all);
    }
    case PC_LIN_EQ: {   // Equality<false>::deduce, right side constant (formula.hpp:676-680)
      const Itv all = lin_project(a, terms, n);
      return lin_embed(a, terms, n, Itv(rhs, rhs), all);
    }
    case PC_LIN_EQ_VAR: {   // Equality<false>::deduce (formula.hpp:672-681): z <- sum, then sum <- z
      int f = a.embed(bvar, lin_project(a, terms, n));
      const Itv z = a.load(bvar);
      const Itv all = lin_project(a, terms, n);
      f |= lin_embed(a, terms, n, z, all);
      return f;
    }
    case PC_EQ: {   // Equality<false>::deduce (formula.hpp:672-681)
      const int x = terms[0].y, y = terms[1].y;
      int f = a.embed(y, a.load(x));
      f |= a.embed(x, a.load(y));
      return f;
    }
    case PC_NEQ: {   // Equality<true>::deduce (formula.hpp:636-670)
      const int x = terms[0].y;
      if(n == 2) {
        const int y = terms[1].y;
        const Itv l = a.load(x);
        if(l.lb == l.ub) {
          Itv r = a.load(y), lo = r, hi = r;
          lo.meet(Itv(b_add(l.lb, 1), LPC_INF));
          hi.meet(Itv(LPC_MINF, b_sub(l.ub, 1)));
          return a.embed(y, fjoin(lo, hi));
        }
        const Itv r = a.load(y);
        if(r.lb == r.ub) {
          Itv l2 = a.load(x), lo = l2, hi = l2;
          lo.meet(Itv(b_add(r.lb, 1), LPC_INF));
          hi.meet(Itv(LPC_MINF, b_sub(r.ub, 1)));
          return a.embed(x, fjoin(lo, hi));
        }
        return 0;
      }
      // x != constant: only the left side can move
      Itv l2 = a.load(x), lo = l2, hi = l2;
      lo.meet(Itv(b_add(rhs, 1), LPC_INF));
      hi.meet(Itv(LPC_MINF, b_sub(rhs, 1)));
      return a.embed(x, fjoin(lo, hi));
    }
    case PC_CLAUSE: {   // nested Disjunction::deduce = unit propagation (formula.hpp:346-350)
      int first = -1;
      bool rest_refuted = true;
      for(int i = 0; i < n; ++i) {
        const int2 t = terms[i];
        const bool refuted = lit_ask(t.x > 0, a.load(t.y));   // nask of a literal = ask of its negation
        if(first < 0) { if(!refuted) first = i; }
        else rest_refuted &= refuted;
      }
      if(first < 0) first = n - 1;   // every literal refuted: the innermost level deduces the last one -> bot
      else if(!rest_refuted) return 0;
      return lit_deduce(a, terms[first].x < 0, terms[first].y);
    }
    case PC_ABS_EQ: {   // Equality(Abs(x), y): y <- |x| (terms.hpp:107-110), then x <- hull(y, -y) (:112-114)
      const int x = terms[0].y, y = terms[1].y;
      const Itv xv = a.load(x);
      Itv ax = xv;
      if(!xv.is_bot()) {
        if(xv.lb >= 0) ax = xv;
        else if(xv.ub <= 0) ax = Itv(b_neg(xv.ub), b_neg(xv.lb));
        else ax = Itv(0, max(b_neg(xv.lb), xv.ub));
      }
      int f = a.embed(y, ax);
      const Itv r = a.load(y);
      f |= a.embed(x, fjoin(r, Itv(b_neg(r.ub), b_neg(r.lb))));
      return f;
    }
    case PC_TREE: return pc_tree_deduce(a, reinterpret_cast<const int*>(terms));
    default: return 0;
  }
}

// PC::ask(i) (pc.hpp:661-663) for one flat propagator.
template <class Acc>
LPC_HD bool pc_ask(const Acc& a, const int4 h, const int2* terms) {
  const int kind = h.x & 0xff, n = h.x >> 8, rhs = h.z, bvar = h.w;
  switch(kind) {
    case PC_LIN_LE: return lin_project(a, terms, n).ub <= rhs;                       // formula.hpp:769
    case PC_REIF_LIN_LE: {                                                           // formula.hpp:408-412
      const Itv b = a.load(bvar), all = lin_project(a, terms, n);
      return (lit_ask(false, b) && all.ub <= rhs) || (lit_ask(true, b) && all.lb > rhs);
    }
    case PC_LIN_GE: return rhs <= lin_project(a, terms, n).lb;                       // formula.hpp:769
    case PC_LIN_GT: return lin_project(a, terms, n).lb > rhs;                        // formula.hpp:766
    case PC_LIN_EQ: { const Itv all = lin_project(a, terms, n); return all.lb == rhs && all.ub == rhs; }   // formula.hpp:629
    case PC_LIN_EQ_VAR: {
      const Itv all = lin_project(a, terms, n), z = a.load(bvar);
      return ((all.is_bot() && z.is_bot()) || (all.lb == z.lb && all.ub == z.ub)) && all.lb == all.ub;
    }
    case PC_EQ: {                                                                    // formula.hpp:629
      const Itv l = a.load(terms[0].y), r = a.load(terms[1].y);
      return ((l.is_bot() && r.is_bot()) || (l.lb == r.lb && l.ub == r.ub)) && l.lb == l.ub;
    }
    case PC_NEQ: {                                                                   // formula.hpp:626
      Itv l = a.load(terms[0].y);
      const Itv r = n == 2 ? a.load(terms[1].y) : Itv(rhs, rhs);
      l.meet(r);
      return l.is_bot();
    }
    case PC_CLAUSE: {                                                                // formula.hpp:338-340
      bool any = false;
      for(int i = 0; i < n; ++i) any |= lit_ask(terms[i].x < 0, a.load(terms[i].y));
      return any;
    }
    case PC_ABS_EQ: {
      const Itv xv = a.load(terms[0].y), r = a.load(terms[1].y);
      Itv ax = xv;
      if(!xv.is_bot()) {
        if(xv.lb >= 0) ax = xv;
        else if(xv.ub <= 0) ax = Itv(b_neg(xv.ub), b_neg(xv.lb));
        else ax = Itv(0, max(b_neg(xv.lb), xv.ub));
      }
      return ((ax.is_bot() && r.is_bot()) || (ax.lb == r.lb && ax.ub == r.ub)) && ax.lb == ax.ub;
    }
    case PC_TREE: return pc_tree_ask(a, reinterpret_cast<const int*>(terms));
    default: return true;
  }
}

// ---- the same propagators over a VStore<NBitset<64>> (tests/pc_bitset_test.cpp:23-25) ----------------------------
// One uint64 per variable: bit 0 = "some value <= -1", bit i (1..62) = value i - 1, bit 63 = "some value >= 62"; meet =
// AND, join = OR, bot = 0, complement = NOT (lala-core nbitset.hpp, un-vendored; pinned by pc_bitset_test.cpp). Only the
// four shapes those tests pin have a flat bitset rule (EQ, NEQ, CLAUSE, ABS_EQ); everything else runs through the tree
// interpreter over the NBitset universe (pc_tree.cuh: arithmetic through the interval hull, unpinned upstream). `BAcc`: `load(v)` -> bits, `embed(v, bits)` -> bit0 =
// changed, bit1 = became empty.
typedef unsigned long long u64;
LPC_HD int nb_ctz(u64 b) {
#ifdef __CUDA_ARCH__
  return __ffsll((long long)b) - 1;
#else
  return __builtin_ctzll(b);
#endif
}
LPC_HD int nb_clz(u64 b) {
#ifdef __CUDA_ARCH__
  return __clzll((long long)b);
#else
  return __builtin_clzll(b);
#endif
}
// NBitset(lb, ub)
LPC_HD u64 nb_range(int l, int u) {
  if(l > u) return 0;
  const int from = l < 0 ? 0 : (l >= 62 ? 63 : l + 1), to = u < 0 ? 0 : (u >= 62 ? 63 : u + 1);
  return (~0ull << from) & (~0ull >> (63 - to));
}
// [lb(), ub()] of a non-empty set; the empty set gives the empty interval
LPC_HD Itv nb_itv(u64 b) {
  if(b == 0) return itv_bot();
  return Itv((b & 1) ? LPC_MINF : nb_ctz(b) - 1, (b >> 63) ? LPC_INF : 62 - nb_clz(b));
}
LPC_HD bool nb_singleton(u64 b) { return b != 0 && (b & (b - 1)) == 0 && !(b & 1) && !(b >> 63); }   // lb() == ub()
LPC_HD u64 nb_abs(u64 b) {   // project(ABS) through the interval (terms.hpp:107-110)
  if(b == 0) return 0;
  const Itv x = nb_itv(b);
  if(x.lb >= 0) return nb_range(x.lb, x.ub);
  if(x.ub <= 0) return nb_range(b_neg(x.ub), b_neg(x.lb));
  return nb_range(0, max(b_neg(x.lb), x.ub));
}
LPC_HD u64 nb_neg(u64 b) {   // project_fun(NEG)
  if(b == 0) return 0;
  const Itv x = nb_itv(b);
  return nb_range(b_neg(x.ub), b_neg(x.lb));
}
LPC_HD bool nb_lit_ask(bool neg, u64 b) { return neg ? (b & ~2ull) == 0 : !((b >> 1) & 1); }   // formula.hpp:100-110

#ifndef LPC_HOST_HARNESS
// VStore<NBitset<64>> in global memory: one uint64 per variable, joins by atomicAnd.
struct GlobalBitAcc {
  typedef UNb Univ;
  u64* s;
  mutable int seen_bot;
  int touched;   // as in GlobalAcc
  __device__ __forceinline__ u64 load(int v) const {
    const u64 d = __ldcg(&s[v]);
    seen_bot |= d == 0;
    return d;
  }
  __device__ __forceinline__ int embed(int v, u64 u) {
    const u64 old = __ldcg(&s[v]);
    if(old == 0) return 2;
    const u64 nw = old & u;
    if(nw == old) return 0;
    atomicAnd(&s[v], u);
    touched |= 1;
    return nw == 0 ? 3 : 1;
  }
};
__device__ int pc_tree_deduce_global_bits(GlobalBitAcc& a, const int* words);
__device__ bool pc_tree_ask_global_bits(const GlobalBitAcc& a, const int* words);
__device__ __forceinline__ int pc_tree_deduce(GlobalBitAcc& a, const int* words) { return pc_tree_deduce_global_bits(a, words); }
__device__ __forceinline__ bool pc_tree_ask(const GlobalBitAcc& a, const int* words) { return pc_tree_ask_global_bits(a, words); }
#endif

template <class BAcc>
LPC_HD int pc_deduce_bits(BAcc& a, const int4 h, const int2* terms) {
  const int kind = h.x & 0xff, n = h.x >> 8, rhs = h.z;
  switch(kind) {
    case PC_EQ: {   // Equality<false>::deduce (formula.hpp:672-681)
      const int x = terms[0].y, y = terms[1].y;
      int f = a.embed(y, a.load(x));
      f |= a.embed(x, a.load(y));
      return f;
    }
    case PC_NEQ: {   // Equality<true>::deduce, complemented universe (formula.hpp:640-644, 656-660)
      const int x = terms[0].y;
      if(n == 2) {
        const int y = terms[1].y;
        const u64 l = a.load(x);
        if(nb_singleton(l)) return a.embed(y, ~l);
        const u64 r = a.load(y);
        if(nb_singleton(r)) return a.embed(x, ~r);
        return 0;
      }
      const u64 r = nb_range(rhs, rhs);            // Constant::project into the universe
      if(nb_singleton(r)) return a.embed(x, ~r);
      return 0;
    }
    case PC_CLAUSE: {   // nested Disjunction::deduce = unit propagation (formula.hpp:346-350)
      int first = -1;
      bool rest_refuted = true;
      for(int i = 0; i < n; ++i) {
        const int2 t = terms[i];
        const bool refuted = nb_lit_ask(t.x > 0, a.load(t.y));
        if(first < 0) { if(!refuted) first = i; }
        else rest_refuted &= refuted;
      }
      if(first < 0) first = n - 1;
      else if(!rest_refuted) return 0;
      return a.embed(terms[first].y, terms[first].x < 0 ? 2ull : 4ull);   // eq_zero = {0}, eq_one = {1}
    }
    case PC_ABS_EQ: {   // Equality(Abs(x), y) (terms.hpp:104-119)
      const int x = terms[0].y, y = terms[1].y;
      int f = a.embed(y, nb_abs(a.load(x)));
      const u64 r = a.load(y);
      f |= a.embed(x, r | nb_neg(r));
      return f;
    }
    // every other shape - sums included - is walked as a tree over the NBitset universe (pc_tree.cuh, struct UNb); the table
    // builder hands the linear kinds over in that form (pc_linear_tree_words)
    case PC_TREE: return pc_tree_deduce(a, reinterpret_cast<const int*>(terms));
    default: return 0;
  }
}

template <class BAcc>
LPC_HD bool pc_ask_bits(const BAcc& a, const int4 h, const int2* terms) {
  const int kind = h.x & 0xff, n = h.x >> 8, rhs = h.z;
  switch(kind) {
    case PC_EQ: { const u64 l = a.load(terms[0].y), r = a.load(terms[1].y); return l == r && nb_singleton(l); }   // formula.hpp:629
    case PC_NEQ: { const u64 l = a.load(terms[0].y), r = n == 2 ? a.load(terms[1].y) : nb_range(rhs, rhs); return (l & r) == 0; }
    case PC_CLAUSE: {
      bool any = false;
      for(int i = 0; i < n; ++i) any |= nb_lit_ask(terms[i].x < 0, a.load(terms[i].y));
      return any;
    }
    case PC_ABS_EQ: { const u64 ax = nb_abs(a.load(terms[0].y)), r = a.load(terms[1].y); return ax == r && nb_singleton(ax); }
    case PC_TREE: return pc_tree_ask(a, reinterpret_cast<const int*>(terms));
    default: return true;
  }
}

} // namespace lpc

#ifdef LPC_HOST_HARNESS
#include "pc_tree.cuh"
#endif
