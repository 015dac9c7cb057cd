// pc_tree.cuh — the general PC propagator: a prefix-encoded formula / term tree interpreted in registers.
//
// The flat kinds of pc_device.cuh cover the shapes that make up large models (sums, reified sums, clauses, =, !=, abs).
// Everything else pc::Formula / pc::Term can hold keeps its tree, here as a prefix-encoded int32 stream that lives in
// the term array of the table (kind LPC_PC_TREE, include/lpc_pc.h), so PC::deduce(i) and PC::ask(i) have a device rule
// for EVERY propagator the reference's interpreter can build from these node types (lala-pc include/lala/):
//   formulas  VariableLiteral<neg>        formula.hpp:80-167      Conjunction            formula.hpp:241-308
//             Disjunction                 formula.hpp:310-378     Biconditional          formula.hpp:380-451
//             Implication                 formula.hpp:453-516     ExclusiveDisjunction   formula.hpp:518-587
//             Equality<neg>               formula.hpp:589-721     Inequality<neg>        formula.hpp:727-845
//   terms     Constant, Variable          terms.hpp:18-85         Unary<Neg | Abs>       terms.hpp:87-175
//             Binary<Add|Sub|Mul|Min|Max> terms.hpp:177-262, 301-434                     Nary<Add|Mul>  terms.hpp:436-526
//             Binary<TDiv|FDiv|CDiv|EDiv> terms.hpp:264-299       AbstractElement        formula.hpp:14-77
// The walk is the reference's: `project` folds a term bottom-up, `embed` pushes an interval down through the residuals
// of each group, formulas dispatch ask / nask / deduce / contradeduce with the `negated` flag folded in.
//
// No recursion on the device: every function is a template over the REMAINING depth and calls the level below, so the
// call graph is static and ptxas sizes the stack exactly. The table builder (lpc_pc_table_create) checks a stream
// against PC_TREE_TERM_DEPTH / PC_TREE_FORM_DEPTH and refuses deeper trees.
//
// Bound arithmetic: the same restatement of lala-core's Interval::project as pc_device.cuh (b_add, b_mul, b_ediv ...).
// Pinned by tests/pc_test.cpp through the CPU checker for what its goldens exercise; MIN / MAX / MUL of two variables
// by MinConstraint1-3, MaxConstraint1-3 and IntTimes1-6.
#pragma once
// (included by pc_tree.cu, by the table builder for tree_check, and - on the host harness - at the end of pc_device.cuh)
#include "pc_device.cuh"
#include <vector>
namespace lpc {

enum PcTok : int { T_CONST = 1, T_VAR = 2, T_NEG = 3, T_ABS = 4, T_ADD = 5, T_SUB = 6, T_MUL = 7, T_NARY_ADD = 8,
                   T_MIN = 9, T_MAX = 10, T_TDIV = 11, T_FDIV = 12, T_CDIV = 13, T_EDIV = 14, T_NARY_MUL = 15,
                   F_VARLIT = 20, F_NVARLIT = 21, F_LEQ = 22, F_GT = 23, F_EQ = 24, F_NEQ = 25, F_AND = 26, F_OR = 27,
                   F_EQUIV = 28, F_IMPLY = 29, F_XOR = 30, F_AE = 31, F_TRUE = 32, F_FALSE = 33 };
// F_TRUE / F_FALSE: the constant formulas (formula.hpp:169-239), no operand; False::deduce / True::contradeduce send the
// store to bot (a.meet_bot()), here by emptying variable 0.
// F_AE op var k: AbstractElement over the store, `var op k` with op = 0 <=, 1 >=, 2 =, 3 != (formula.hpp:14-77)
enum PcAeOp : int { AE_LEQ = 0, AE_GEQ = 1, AE_EQ = 2, AE_NEQ = 3 };

constexpr int PC_TREE_TERM_DEPTH = 8;   // height of the tallest term  (x + y = 2, (2*x + y) + 3*z = 4)
constexpr int PC_TREE_FORM_DEPTH = 6;   // height of the connective nest above a comparison (b <=> (p /\ q) = 3)

#ifdef LPC_HOST_HARNESS
#define LPC_NI __host__ __device__ __noinline__
#else
#define LPC_NI __device__ __noinline__
#endif

__host__ __device__ inline bool tok_is_div(int k) { return k >= T_TDIV && k <= T_EDIV; }
__host__ __device__ inline bool tok_is_binary_term(int k) { return k == T_ADD || k == T_SUB || k == T_MUL || k == T_MIN || k == T_MAX || tok_is_div(k); }

// The first word after the term that starts at p (iterative: `pending` subterms still to be read).
__host__ __device__ inline const int* tree_skip_term(const int* p) {
  int pending = 1;
  while(pending > 0) {
    const int k = *p++;
    if(k == T_CONST || k == T_VAR) { ++p; --pending; }
    else if(k == T_NEG || k == T_ABS) { }
    else if(k == T_NARY_ADD || k == T_NARY_MUL) { pending += *p++ - 1; }
    else ++pending;   // binary
  }
  return p;
}
__host__ __device__ inline const int* tree_skip_formula(const int* p) {
  int pending = 1;
  while(pending > 0) {
    const int k = *p++;
    if(k == F_VARLIT || k == F_NVARLIT) { ++p; --pending; }
    else if(k == F_AE) { p += 3; --pending; }
    else if(k == F_TRUE || k == F_FALSE) { --pending; }
    else if(k >= F_LEQ && k <= F_NEQ) { p = tree_skip_term(tree_skip_term(p)); --pending; }
    else ++pending;   // binary connective
  }
  return p;
}

// Host-side check of one stream (table builder): well formed, variables in range, depths within the limits.
// Returns the number of words the formula occupies, or -1.
struct TreeCheck {
  const int* w; int n; int nvars; bool ok;
  // `left` = levels this subtree may still use: the recursion stops where the device interpreter's static depth does,
  // so a hostile stream (a long chain of negations) cannot overflow the host stack.
  int term(int& i, int left) {   // returns the height
    if(i >= n || left <= 0) { ok = false; return 0; }
    const int k = w[i++];
    if(k == T_CONST) { if(i >= n) { ok = false; return 0; } ++i; return 1; }
    if(k == T_VAR) { if(i >= n || w[i] < 0 || w[i] >= nvars) { ok = false; return 0; } ++i; return 1; }
    if(k == T_NEG || k == T_ABS) { const int a = term(i, left - 1); return ok ? 1 + a : 0; }
    if(tok_is_binary_term(k)) {
      const int a = term(i, left - 1); if(!ok) return 0;
      const int b = term(i, left - 1); if(!ok) return 0;
      return 1 + (a > b ? a : b);
    }
    if(k == T_NARY_ADD || k == T_NARY_MUL) {
      if(i >= n || w[i] < 2 || w[i] > n) { ok = false; return 0; }
      const int m = w[i++];
      int h = 0;
      for(int j = 0; j < m && ok; ++j) { const int a = term(i, left - 1); h = a > h ? a : h; }
      return ok ? 1 + h : 0;
    }
    ok = false;
    return 0;
  }
  int formula(int& i, int left) {
    if(i >= n || left <= 0) { ok = false; return 0; }
    const int k = w[i++];
    if(k == F_VARLIT || k == F_NVARLIT) { if(i >= n || w[i] < 0 || w[i] >= nvars) { ok = false; return 0; } ++i; return 1; }
    if(k == F_TRUE) return 1;
    if(k == F_FALSE) { if(nvars < 1) ok = false; return 1; }   // needs a variable to empty
    if(k == F_AE) { if(i + 2 >= n || w[i] < AE_LEQ || w[i] > AE_NEQ || w[i + 1] < 0 || w[i + 1] >= nvars) { ok = false; return 0; } i += 3; return 1; }
    if(k >= F_LEQ && k <= F_NEQ) {
      term(i, PC_TREE_TERM_DEPTH); if(!ok) return 0;
      term(i, PC_TREE_TERM_DEPTH); if(!ok) return 0;
      return 1;
    }
    if(k >= F_AND && k <= F_XOR) {
      const int a = formula(i, left - 1); if(!ok) return 0;
      const int b = formula(i, left - 1); if(!ok) return 0;
      return 1 + (a > b ? a : b);
    }
    ok = false;
    return 0;
  }
};
inline int tree_check(const int* w, int n, int nvars) {
  TreeCheck c{w, n, nvars, true};
  int i = 0;
  const int h = c.formula(i, PC_TREE_FORM_DEPTH);
  if(!c.ok || h > PC_TREE_FORM_DEPTH) return -1;
  for(int j = i; j < n; ++j) if(w[j] != 0) return -1;   // only zero padding after the formula
  return i;
}

// A flat linear propagator (include/lpc_pc.h) as the formula stream of the tree it stands for - the shape the
// reference's interpreter builds (pc.hpp:277-296): a coefficient 1 is the bare variable, any other c * x is
// Binary<Mul>(Constant, Variable); one term stands alone, two are Binary<Add>, three or more Nary<Add>. Used by the table
// builder for stores whose universe has no lane-tile rule for sums (NBitset).
inline void pc_linear_tree_words(int kind, const int2* terms, int n, int rhs, int bvar, std::vector<int>& w) {
  auto lhs = [&]() {
    if(n == 2) w.push_back(T_ADD);
    else if(n > 2) { w.push_back(T_NARY_ADD); w.push_back(n); }
    for(int i = 0; i < n; ++i) {
      if(terms[i].x != 1) { w.push_back(T_MUL); w.push_back(T_CONST); w.push_back(terms[i].x); }
      w.push_back(T_VAR); w.push_back(terms[i].y);
    }
  };
  auto cst = [&]() { w.push_back(T_CONST); w.push_back(rhs); };
  switch(kind) {
    case PC_LIN_LE: w.push_back(F_LEQ); lhs(); cst(); break;
    case PC_LIN_GE: w.push_back(F_LEQ); cst(); lhs(); break;
    case PC_LIN_GT: w.push_back(F_GT); lhs(); cst(); break;
    case PC_LIN_EQ: w.push_back(F_EQ); lhs(); cst(); break;
    case PC_LIN_EQ_VAR: w.push_back(F_EQ); lhs(); w.push_back(T_VAR); w.push_back(bvar); break;
    default: w.push_back(F_EQUIV); w.push_back(F_VARLIT); w.push_back(bvar); w.push_back(F_LEQ); lhs(); cst(); break;   // PC_REIF_LIN_LE
  }
  if(w.size() % 2) w.push_back(0);
}

// The table as a bitset store sees it: the four flat kinds with a bitset rule stay, every linear propagator becomes an
// LPC_PC_TREE propagator whose stream is appended to the term array. props: n x {kind, first_term, n_terms, rhs, bvar}.
inline void pc_bits_view(const int* props, long long n, const int2* terms, long long n_terms, std::vector<int>& oprops,
                         std::vector<int2>& oterms) {
  oprops.assign(props, props + 5 * n);
  oterms.assign(terms, terms + n_terms);
  std::vector<int> w;
  for(long long i = 0; i < n; ++i) {
    int* p = oprops.data() + 5 * i;
    if(!pc_is_linear(p[0])) continue;
    w.clear();
    pc_linear_tree_words(p[0], terms + p[1], p[2], p[3], p[4], w);
    p[0] = PC_TREE; p[1] = (int)oterms.size(); p[2] = (int)(w.size() / 2); p[3] = 0; p[4] = -1;
    for(size_t k = 0; k < w.size(); k += 2) oterms.push_back(make_int2(w[k], w[k + 1]));
  }
}

// ---- interval helpers (lala-core Interval::project(Sig, ...), restated; see the header of pc_device.cuh) -------------
LPC_HD Itv tr_neg(const Itv& a) { return Itv(b_neg(a.ub), b_neg(a.lb)); }
LPC_HD Itv tr_add(const Itv& a, const Itv& b) { return Itv(b_add(a.lb, b.lb), b_add(a.ub, b.ub)); }
LPC_HD Itv tr_sub(const Itv& a, const Itv& b) { return Itv(b_sub(a.lb, b.ub), b_sub(a.ub, b.lb)); }
LPC_HD Itv tr_abs(const Itv& a) {
  if(a.is_bot() || a.lb >= 0) return a;
  if(a.ub <= 0) return tr_neg(a);
  return Itv(0, max(b_neg(a.lb), a.ub));
}
LPC_HD Itv tr_hull4(int t1, int t2, int t3, int t4) { return Itv(min(min(t1, t2), min(t3, t4)), max(max(t1, t2), max(t3, t4))); }
LPC_HD Itv tr_mul(const Itv& a, const Itv& b) {
  if(a.is_bot() || b.is_bot()) return itv_bot();
  return tr_hull4(b_mul(a.lb, b.lb), b_mul(a.lb, b.ub), b_mul(a.ub, b.lb), b_mul(a.ub, b.ub));
}
LPC_HD int tr_ediv1(int a, int b) {   // Euclidean quotient of two bounds, b != 0
  if(b_inf(b)) return b_inf(a) ? (((a < 0) != (b < 0)) ? LPC_MINF : LPC_INF) : 0;
  return b_ediv(a, b);
}
// a divisor holding 0 loses it: hull over its negative and its positive part (0 at an endpoint is pinned by IntTimes2,
// pc_test.cpp:626-636; a straddling divisor and {0} -> empty are not pinned by any reference test)
LPC_HD Itv tr_ediv(const Itv& a, const Itv& b) {
  if(a.is_bot() || b.is_bot()) return itv_bot();
  Itv r = itv_bot();
  if(b.lb < 0) {
    const int bu = min(b.ub, -1);
    r = fjoin(r, tr_hull4(tr_ediv1(a.lb, b.lb), tr_ediv1(a.lb, bu), tr_ediv1(a.ub, b.lb), tr_ediv1(a.ub, bu)));
  }
  if(b.ub > 0) {
    const int bl = max(b.lb, 1);
    r = fjoin(r, tr_hull4(tr_ediv1(a.lb, bl), tr_ediv1(a.lb, b.ub), tr_ediv1(a.ub, bl), tr_ediv1(a.ub, b.ub)));
  }
  return r;
}
LPC_HD Itv tr_min(const Itv& a, const Itv& b) {
  if(a.is_bot() || b.is_bot()) return itv_bot();
  return Itv(min(a.lb, b.lb), min(a.ub, b.ub));
}
LPC_HD Itv tr_max(const Itv& a, const Itv& b) {
  if(a.is_bot() || b.is_bot()) return itv_bot();
  return Itv(max(a.lb, b.lb), max(a.ub, b.ub));
}
// project(TDIV | FDIV | CDIV | EDIV): the corner quotients with the operator's rounding, 0 cut out of the divisor.
// Pinned on 0/1 domains by IntDiv1-2 (pc_test.cpp:678-698); not pinned by any reference test beyond that.
LPC_HD int tr_div1(int k, int a, int b) {   // b != 0
  if(k == T_EDIV) return tr_ediv1(a, b);
  if(b_inf(a)) return b > 0 ? a : b_neg(a);
  if(b_inf(b)) return 0;
  long long q = (long long)a / b;
  const long long r = (long long)a % b;
  if(k == T_FDIV && r != 0 && ((r < 0) != (b < 0))) --q;
  if(k == T_CDIV && r != 0 && ((r < 0) == (b < 0))) ++q;
  return b_clamp(q);
}
LPC_HD Itv tr_div(int k, const Itv& a, const Itv& b) {
  if(a.is_bot() || b.is_bot()) return itv_bot();
  Itv r = itv_bot();
  if(b.lb < 0) {
    const int bu = min(b.ub, -1);
    r = fjoin(r, tr_hull4(tr_div1(k, a.lb, b.lb), tr_div1(k, a.lb, bu), tr_div1(k, a.ub, b.lb), tr_div1(k, a.ub, bu)));
  }
  if(b.ub > 0) {
    const int bl = max(b.lb, 1);
    r = fjoin(r, tr_hull4(tr_div1(k, a.lb, bl), tr_div1(k, a.lb, b.ub), tr_div1(k, a.ub, bl), tr_div1(k, a.ub, b.ub)));
  }
  return r;
}
// GroupDiv::left_residual (terms.hpp:272-277): x in u * y, its upper bound joined with y.ub - 1
LPC_HD Itv tr_div_left(const Itv& u, const Itv& b) {
  Itv r = tr_mul(u, b);
  if(!r.is_bot()) r.ub = max(r.ub, b_sub(b.ub, 1));
  return r;
}
// GroupDiv::right_residual (terms.hpp:279-294) on r = the divisor's current value: 0 leaves its ends, then
// y in x / u unless x holds 0 or u = {0}
LPC_HD Itv tr_div_right(int k, const Itv& u, const Itv& b, Itv r) {
  if(r.lb == 0) r.meet(Itv(1, LPC_INF));
  if(r.ub == 0) r.meet(Itv(LPC_MINF, -1));
  if(!contains0(b) && !(u.lb == 0 && u.ub == 0)) r.meet(tr_div(k, b, u));
  return r;
}
LPC_HD bool tr_same(const Itv& a, const Itv& b) { return (a.is_bot() && b.is_bot()) || (a.lb == b.lb && a.ub == b.ub); }

// ---- universes ----------------------------------------------------------------------------------------------------------
// The walk below is written once over a universe U (the reference's template parameter of pc::Term / pc::Formula): the
// value type V of a variable's domain and of every intermediate result, with the operations Interval / NBitset offer.
// An accessor names its universe (Acc::Univ); load(v) returns a V, embed(v, V) joins one.
//  * UItv: Interval<ZLB>, the operations above.
//  * UNb: NBitset<64> as one uint64 (pc_device.cuh). lala-core's NBitset::project is un-vendored and no reference test
//    pins it beyond IntAbs1: every arithmetic operation goes through the interval hull of its operands and
//    back into the universe (so results beyond [-1, 62] fold into the two open-ended bits), the set operations - meet,
//    join, complement, inclusion - are bitwise. The CPU checker of the test-suite carries the same
//    reading; PARITY UNPINNED against upstream for anything but =, !=, clauses and abs.
struct UItv {
  typedef Itv V;
  static constexpr bool complemented = false;
  static LPC_HD V range(int l, int u) { return Itv(l, u); }
  static LPC_HD V top() { return itv_top(); }
  static LPC_HD V bot() { return itv_bot(); }
  static LPC_HD int lo(const V& a) { return a.lb; }
  static LPC_HD int hi(const V& a) { return a.ub; }
  static LPC_HD bool is_bot(const V& a) { return a.is_bot(); }
  static LPC_HD void meet(V& a, const V& b) { a.meet(b); }
  static LPC_HD V join(const V& a, const V& b) { return fjoin(a, b); }
  static LPC_HD V complement(const V& a) { return a; }   // never called: complemented == false
  static LPC_HD bool same(const V& a, const V& b) { return tr_same(a, b); }
  static LPC_HD bool has0(const V& a) { return contains0(a); }
  static LPC_HD bool sub_of(const V& a, int l, int u) { return a.is_bot() || (a.lb >= l && a.ub <= u); }
  static LPC_HD V neg(const V& a) { return tr_neg(a); }
  static LPC_HD V abs(const V& a) { return tr_abs(a); }
  static LPC_HD V add(const V& a, const V& b) { return tr_add(a, b); }
  static LPC_HD V sub(const V& a, const V& b) { return tr_sub(a, b); }
  static LPC_HD V mul(const V& a, const V& b) { return tr_mul(a, b); }
  static LPC_HD V ediv(const V& a, const V& b) { return tr_ediv(a, b); }
  static LPC_HD V vmin(const V& a, const V& b) { return tr_min(a, b); }
  static LPC_HD V vmax(const V& a, const V& b) { return tr_max(a, b); }
  static LPC_HD V div(int k, const V& a, const V& b) { return tr_div(k, a, b); }
  static LPC_HD V div_left(const V& u, const V& b) { return tr_div_left(u, b); }
  static LPC_HD V div_right(int k, const V& u, const V& b, const V& r) { return tr_div_right(k, u, b, r); }
  static LPC_HD V only_lb(const V& a) { return Itv(a.lb, LPC_INF); }
  static LPC_HD V only_ub(const V& a) { return Itv(LPC_MINF, a.ub); }
  // GroupAdd::rev_op's operand (terms.hpp:190-194): the crossed interval [-lb, -ub], so that all + it = all - t bound by bound
  static LPC_HD V additive_inverse(const V& t) { return Itv(b_neg(t.lb), b_neg(t.ub)); }
};
struct UNb {
  typedef u64 V;
  static constexpr bool complemented = true;
  static LPC_HD V from(const Itv& i) { return nb_range(i.lb, i.ub); }
  static LPC_HD Itv itv(V a) { return nb_itv(a); }
  static LPC_HD V range(int l, int u) { return nb_range(l, u); }
  static LPC_HD V top() { return ~0ull; }
  static LPC_HD V bot() { return 0ull; }
  static LPC_HD int lo(V a) { return nb_itv(a).lb; }
  static LPC_HD int hi(V a) { return nb_itv(a).ub; }
  static LPC_HD bool is_bot(V a) { return a == 0; }
  static LPC_HD void meet(V& a, V b) { a &= b; }
  static LPC_HD V join(V a, V b) { return a | b; }
  static LPC_HD V complement(V a) { return ~a; }
  static LPC_HD bool same(V a, V b) { return a == b; }
  static LPC_HD bool has0(V a) { return (a >> 1) & 1; }
  static LPC_HD bool sub_of(V a, int l, int u) { return (a & ~nb_range(l, u)) == 0; }
  static LPC_HD V neg(V a) { return a == 0 ? a : from(tr_neg(itv(a))); }
  static LPC_HD V abs(V a) { return a == 0 ? a : from(tr_abs(itv(a))); }
  static LPC_HD V add(V a, V b) { return (a == 0 || b == 0) ? 0 : from(tr_add(itv(a), itv(b))); }
  static LPC_HD V sub(V a, V b) { return (a == 0 || b == 0) ? 0 : from(tr_sub(itv(a), itv(b))); }
  static LPC_HD V mul(V a, V b) { return (a == 0 || b == 0) ? 0 : from(tr_mul(itv(a), itv(b))); }
  static LPC_HD V ediv(V a, V b) { return (a == 0 || b == 0) ? 0 : from(tr_ediv(itv(a), itv(b))); }
  static LPC_HD V vmin(V a, V b) { return (a == 0 || b == 0) ? 0 : from(tr_min(itv(a), itv(b))); }
  static LPC_HD V vmax(V a, V b) { return (a == 0 || b == 0) ? 0 : from(tr_max(itv(a), itv(b))); }
  static LPC_HD V div(int k, V a, V b) { return (a == 0 || b == 0) ? 0 : from(tr_div(k, itv(a), itv(b))); }
  static LPC_HD V div_left(V u, V b) { return (u == 0 || b == 0) ? 0 : from(tr_div_left(itv(u), itv(b))); }
  static LPC_HD V div_right(int k, V u, V b, V r) { return r == 0 ? r : from(tr_div_right(k, itv(u), itv(b), itv(r))); }
  static LPC_HD V only_lb(V a) { return nb_range(lo(a), LPC_INF); }
  static LPC_HD V only_ub(V a) { return nb_range(LPC_MINF, hi(a)); }
  static LPC_HD V additive_inverse(V t) { return neg(t); }   // a set has no crossed form
};

// G::project (terms.hpp:182, 213, 235, 306)
template <class U> LPC_HD typename U::V tr_group_u(int k, const typename U::V& x, const typename U::V& y) {
  switch(k) {
    case T_ADD: return U::add(x, y);
    case T_SUB: return U::sub(x, y);
    case T_MUL: return U::mul(x, y);
    case T_MIN: return U::vmin(x, y);
    default: return U::vmax(x, y);
  }
}
// G::left_residual(u, b) / right_residual(u, b) met into top (terms.hpp:196-202, 218-224, 249-257, 310-326)
template <class U> LPC_HD typename U::V tr_residual_u(int k, bool right, const typename U::V& u, const typename U::V& b) {
  switch(k) {
    case T_ADD: return U::sub(u, b);
    case T_SUB: return right ? U::sub(b, u) : U::add(u, b);
    case T_MUL: return (U::has0(u) && U::has0(b)) ? U::top() : U::ediv(u, b);
    default: {   // GroupMinMax: disjoint from the other operand -> this operand IS the result; else only one side
      typename U::V m = u;
      U::meet(m, b);
      if(U::is_bot(m)) return u;
      return k == T_MIN ? U::only_lb(u) : U::only_ub(u);
    }
  }
}

// ---- terms --------------------------------------------------------------------------------------------------------------
template <class U, int D> struct TreeTerm {
  typedef typename U::V V;
  // Term::project (terms.hpp:34, 69-71, 148-152, 399-405, 465-478); p is left after the term
  template <class Acc> static LPC_NI V project(const Acc& a, const int*& p) {
    const int k = *p++;
    switch(k) {
      case T_CONST: { const int c = *p++; return U::range(c, c); }
      case T_VAR: return a.load(*p++);
      case T_NEG: return U::neg(TreeTerm<U, D - 1>::project(a, p));
      case T_ABS: return U::abs(TreeTerm<U, D - 1>::project(a, p));
      case T_NARY_ADD: case T_NARY_MUL: {
        const int n = *p++;
        V accu = TreeTerm<U, D - 1>::project(a, p);
        for(int i = 1; i < n; ++i) {
          const V ti = TreeTerm<U, D - 1>::project(a, p);
          accu = k == T_NARY_ADD ? U::add(accu, ti) : U::mul(accu, ti);
        }
        return accu;
      }
      default: {
        const V x = TreeTerm<U, D - 1>::project(a, p);
        const V y = TreeTerm<U, D - 1>::project(a, p);
        return tok_is_div(k) ? U::div(k, x, y) : tr_group_u<U>(k, x, y);
      }
    }
  }
  // Term::embed(u) (terms.hpp:33, 65-67, 142-146, 376-397, 480-499): bit0 = changed, bit1 = a variable became empty
  template <class Acc> static LPC_NI int embed(Acc& a, const int*& p, const V u) {
    const int k = *p++;
    switch(k) {
      case T_CONST: ++p; return 0;
      case T_VAR: return a.embed(*p++, u);
      case T_NEG: return TreeTerm<U, D - 1>::embed(a, p, U::neg(u));
      case T_ABS: return TreeTerm<U, D - 1>::embed(a, p, U::join(u, U::neg(u)));
      case T_NARY_ADD: case T_NARY_MUL: {
        const int n = *p++;
        const int* q = p - 2;
        const V all = project(a, q);   // once, before any operand moves (terms.hpp:486)
        int f = 0;
        const bool absorbed = k == T_NARY_MUL && U::same(all, U::range(0, 0));   // GroupMul::is_absorbing (:239-241, 487)
        for(int i = 0; i < n; ++i) {
          if(absorbed) { p = tree_skip_term(p); continue; }
          const int* s = p;
          const V ti = TreeTerm<U, D - 1>::project(a, s);
          V res;
          if(k == T_NARY_ADD) {
            const V others = U::add(all, U::additive_inverse(ti));   // GroupAdd::rev_op, :190-194
            res = U::sub(u, others);                                 // left_residual, :196-198
          }
          else {
            const V others = U::ediv(all, ti);                                              // rev_op, :244-246
            res = (U::has0(u) && U::has0(others)) ? U::top() : U::ediv(u, others);          // left_residual, :249-253
          }
          f |= TreeTerm<U, D - 1>::embed(a, p, res);
        }
        return f;
      }
      default: {
        const int* px = p;
        const int* py = tree_skip_term(px);
        int f = 0;
        const bool dv = tok_is_div(k);
        if(*px != T_CONST) {
          const int* s = py;
          const V yt = TreeTerm<U, D - 1>::project(a, s);
          s = px;
          f |= TreeTerm<U, D - 1>::embed(a, s, dv ? U::div_left(u, yt) : tr_residual_u<U>(k, false, u, yt));
        }
        p = tree_skip_term(py);
        if(*py != T_CONST) {
          const int* s = px;
          const V xt = TreeTerm<U, D - 1>::project(a, s);   // re-read: x may just have moved
          V res;
          if(dv) { s = py; res = U::div_right(k, u, xt, TreeTerm<U, D - 1>::project(a, s)); }   // the divisor's own value (:389-392)
          else res = tr_residual_u<U>(k, true, u, xt);
          s = py;
          f |= TreeTerm<U, D - 1>::embed(a, s, res);
        }
        return f;
      }
    }
  }
};
template <class U> struct TreeTerm<U, 0> {   // below the checked depth: never reached (tree_check)
  typedef typename U::V V;
  template <class Acc> static LPC_HD V project(const Acc&, const int*& p) { p = tree_skip_term(p); return U::top(); }
  template <class Acc> static LPC_HD int embed(Acc&, const int*& p, const V) { p = tree_skip_term(p); return 0; }
};

// ---- formulas -----------------------------------------------------------------------------------------------------------
// Equality<true>::deduce, one direction: `other` loses the value of the singleton side (formula.hpp:645-652, 661-668);
// a complemented universe embeds the complement of the singleton (:640-644, 656-660)
template <class U, class Acc> LPC_HD int tree_shave(Acc& a, const int* other, const typename U::V& single) {
  typedef TreeTerm<U, PC_TREE_TERM_DEPTH> TreeTop;
  const int* s = other;
  if(U::complemented) return TreeTop::embed(a, s, U::complement(single));
  const typename U::V o = TreeTop::project(a, s);
  typename U::V lo = o, hi = o;
  U::meet(lo, U::range(b_add(U::lo(single), 1), LPC_INF));
  U::meet(hi, U::range(LPC_MINF, b_sub(U::hi(single), 1)));
  s = other;
  return TreeTop::embed(a, s, U::join(lo, hi));
}

// The comparisons (leaves of the connective nest). `neg` already folds the caller's negation into the node's own.
template <class U, class Acc> LPC_NI bool tree_cmp_ask(const Acc& a, int k, bool negated, const int*& p) {
  typedef TreeTerm<U, PC_TREE_TERM_DEPTH> TreeTop;
  typedef typename U::V V;
  const V x = TreeTop::project(a, p);
  const V y = TreeTop::project(a, p);
  if(k == F_LEQ || k == F_GT) {   // formula.hpp:757-771
    const bool neg = (k == F_GT) != negated;
    return neg ? U::lo(x) > U::hi(y) : U::hi(x) <= U::lo(y);
  }
  const bool neg = (k == F_NEQ) != negated;   // formula.hpp:616-631
  if(neg) { V m = x; U::meet(m, y); return U::is_bot(m); }
  return U::same(x, y) && U::lo(x) == U::hi(x);
}
template <class U, class Acc> LPC_NI int tree_cmp_deduce(Acc& a, int k, bool negated, const int*& p) {
  typedef TreeTerm<U, PC_TREE_TERM_DEPTH> TreeTop;
  typedef typename U::V V;
  const int* pl = p;
  const int* pr = tree_skip_term(pl);
  p = tree_skip_term(pr);
  const bool lconst = *pl == T_CONST, rconst = *pr == T_CONST;
  const int* s;
  int f = 0;
  if(k == F_LEQ || k == F_GT) {   // formula.hpp:773-807
    const bool neg = (k == F_GT) != negated;
    if(neg) {   // l > r: l >= (what is left of r above its lower bound).lb, r <= (what is left of l below its upper bound).ub
      if(!lconst) { s = pr; V y = TreeTop::project(a, s); U::meet(y, U::range(b_add(U::lo(y), 1), LPC_INF)); s = pl; f = TreeTop::embed(a, s, U::range(U::lo(y), LPC_INF)); }
      if(!rconst) { s = pl; V x = TreeTop::project(a, s); U::meet(x, U::range(LPC_MINF, b_sub(U::hi(x), 1))); s = pr; f |= TreeTop::embed(a, s, U::range(LPC_MINF, U::hi(x))); }
    }
    else {      // l <= r
      if(!lconst) { s = pr; const V y = TreeTop::project(a, s); s = pl; f |= TreeTop::embed(a, s, U::range(LPC_MINF, U::hi(y))); }
      if(!rconst) {   // formula.hpp:803 ASSIGNS has_changed here: the left side's change bit is lost, its bot bit is not
        s = pl; const V x = TreeTop::project(a, s); s = pr;
        f = (f & 2) | TreeTop::embed(a, s, U::range(U::lo(x), LPC_INF));
      }
    }
    return f;
  }
  const bool neg = (k == F_NEQ) != negated;   // formula.hpp:633-683
  if(neg) {
    if(!rconst) { s = pl; const V x = TreeTop::project(a, s); if(U::lo(x) == U::hi(x)) return tree_shave<U>(a, pr, x); }
    if(!lconst) { s = pr; const V y = TreeTop::project(a, s); if(U::lo(y) == U::hi(y)) return tree_shave<U>(a, pl, y); }
    return 0;
  }
  if(!rconst) { s = pl; const V x = TreeTop::project(a, s); s = pr; f = TreeTop::embed(a, s, x); }
  if(!lconst) { s = pr; const V y = TreeTop::project(a, s); s = pl; f |= TreeTop::embed(a, s, y); }
  return f;
}

// AbstractElement (formula.hpp:14-77): ask / nask = the store's ask of the element / of its negation, deduce /
// contradeduce = the store's deduce of them. Over an interval store `v != c` has an ask (c outside the domain) but no
// tell (AbstractElement3-4, pc_test.cpp:738-764); a complemented universe tells the complement of {c}.
LPC_HD void tree_ae_norm(int& op, int& c, bool negated) {
  if(!negated) return;
  if(op == AE_LEQ) { op = AE_GEQ; c = b_add(c, 1); }
  else if(op == AE_GEQ) { op = AE_LEQ; c = b_sub(c, 1); }
  else op = op == AE_EQ ? AE_NEQ : AE_EQ;
}
template <class U, class Acc> LPC_HD bool tree_ae_ask(const Acc& a, int op, int v, int c, bool negated) {
  tree_ae_norm(op, c, negated);
  const typename U::V d = a.load(v);
  if(op == AE_NEQ) { typename U::V m = d; U::meet(m, U::range(c, c)); return U::is_bot(m); }
  return op == AE_LEQ ? U::sub_of(d, LPC_MINF, c) : op == AE_GEQ ? U::sub_of(d, c, LPC_INF) : U::sub_of(d, c, c);
}
template <class U, class Acc> LPC_HD int tree_ae_tell(Acc& a, int op, int v, int c, bool negated) {
  tree_ae_norm(op, c, negated);
  if(op == AE_NEQ) return U::complemented ? a.embed(v, U::complement(U::range(c, c))) : 0;
  return a.embed(v, op == AE_LEQ ? U::range(LPC_MINF, c) : op == AE_GEQ ? U::range(c, LPC_INF) : U::range(c, c));
}

template <class U, int D> struct TreeForm {
  // ask (negated = false) / nask (negated = true); p is left after the formula
  template <class Acc> static LPC_NI bool ask(const Acc& a, const int*& p, bool negated) {
    const int k = *p++;
    if(k == F_VARLIT || k == F_NVARLIT) {   // formula.hpp:97-110, 126-135
      const bool neg = (k == F_NVARLIT) != negated;
      const typename U::V t = a.load(*p++);
      return neg ? U::sub_of(t, 0, 0) : !U::has0(t);
    }
    if(k == F_AE) { const int op = p[0], v = p[1], c = p[2]; p += 3; return tree_ae_ask<U>(a, op, v, c, negated); }
    // formula.hpp:182-183, 220-221: true is entailed, false is refuted (their a.is_bot() halves only hold on a failed
    // store, where the fixpoint has stopped)
    if(k == F_TRUE || k == F_FALSE) return (k == F_TRUE) != negated;
    if(k >= F_LEQ && k <= F_NEQ) return tree_cmp_ask<U>(a, k, negated, p);
    const int* pf = p;
    const int* pg = tree_skip_formula(pf);
    p = tree_skip_formula(pg);
    const int *s = pf, *t = pg;
    typedef TreeForm<U, D - 1> S;
    switch(k) {
      case F_AND:     // formula.hpp:268-274
        if(negated) { if(S::ask(a, s, true)) return true; return S::ask(a, t, true); }
        if(!S::ask(a, s, false)) return false;
        return S::ask(a, t, false);
      case F_OR:      // formula.hpp:338-344
        if(negated) { if(!S::ask(a, s, true)) return false; return S::ask(a, t, true); }
        if(S::ask(a, s, false)) return true;
        return S::ask(a, t, false);
      case F_IMPLY:   // formula.hpp:479-487: (not f) or g; negated: f and (not g)
        if(negated) { if(!S::ask(a, s, false)) return false; return S::ask(a, t, true); }
        if(S::ask(a, s, true)) return true;
        return S::ask(a, t, false);
      case F_EQUIV: case F_XOR: {   // formula.hpp:408-419, 541-552
        const int *s2 = pf, *t2 = pg;
        const bool fa = S::ask(a, s, false), ga = S::ask(a, t, false), fn = S::ask(a, s2, true), gn = S::ask(a, t2, true);
        if(k == F_EQUIV) return negated ? ((fa && gn) || (fn && ga)) : ((fa && ga) || (fn && gn));
        return negated ? ((fa && ga) || (fa && gn))   // as written at formula.hpp:548-552
                       : ((fa && gn) || (fn && ga));
      }
      default: return false;
    }
  }
  // deduce (negated = false) / contradeduce (negated = true)
  template <class Acc> static LPC_NI int deduce(Acc& a, const int*& p, bool negated) {
    const int k = *p++;
    if(k == F_VARLIT || k == F_NVARLIT) {   // formula.hpp:112-120, 140-149
      const bool neg = (k == F_NVARLIT) != negated;
      return a.embed(*p++, neg ? U::range(0, 0) : U::range(1, 1));
    }
    if(k == F_AE) { const int op = p[0], v = p[1], c = p[2]; p += 3; return tree_ae_tell<U>(a, op, v, c, negated); }
    // formula.hpp:185-193, 222-228: deducing false / contradeducing true is a.meet_bot()
    if(k == F_TRUE || k == F_FALSE) return ((k == F_FALSE) != negated) ? a.embed(0, U::bot()) : 0;
    if(k >= F_LEQ && k <= F_NEQ) return tree_cmp_deduce<U>(a, k, negated, p);
    const int* pf = p;
    const int* pg = tree_skip_formula(pf);
    p = tree_skip_formula(pg);
    typedef TreeForm<U, D - 1> S;
    auto ASK = [&](const int* q, bool n) { return S::ask(a, q, n); };
    auto DED = [&](const int* q, bool n) { return S::deduce(a, q, n); };
    switch(k) {
      case F_AND:     // formula.hpp:276-286
        if(!negated) { int f = DED(pf, false); f |= DED(pg, false); return f; }
        if(ASK(pf, false)) return DED(pg, true);
        if(ASK(pg, false)) return DED(pf, true);
        return 0;
      case F_OR:      // formula.hpp:346-356
        if(negated) { int f = DED(pf, true); f |= DED(pg, true); return f; }
        if(ASK(pf, true)) return DED(pg, false);
        if(ASK(pg, true)) return DED(pf, false);
        return 0;
      case F_IMPLY:   // formula.hpp:489-499
        if(ASK(pf, false)) return DED(pg, negated);
        if(ASK(pg, true)) return DED(pf, !negated);
        return 0;
      case F_EQUIV: case F_XOR: {   // formula.hpp:421-435, 554-568: xor is the biconditional with the conclusions swapped
        const bool flip = (k == F_XOR) != negated;
        if(ASK(pf, false)) return DED(pg, flip);
        if(ASK(pf, true)) return DED(pg, !flip);
        if(ASK(pg, false)) return DED(pf, flip);
        if(ASK(pg, true)) return DED(pf, !flip);
        return 0;
      }
      default: return 0;
    }
  }
};
template <class U> struct TreeForm<U, 0> {
  template <class Acc> static LPC_HD bool ask(const Acc&, const int*& p, bool) { p = tree_skip_formula(p); return false; }
  template <class Acc> static LPC_HD int deduce(Acc&, const int*& p, bool) { p = tree_skip_formula(p); return 0; }
};

// PC::deduce(i) / PC::ask(i) (pc.hpp:661-680) for one LPC_PC_TREE propagator whose stream starts at `words`, over the
// universe of the accessor.
template <class Acc> LPC_HD int pc_tree_deduce_impl(Acc& a, const int* words) {
  const int* p = words;
  return TreeForm<typename Acc::Univ, PC_TREE_FORM_DEPTH>::deduce(a, p, false);
}
template <class Acc> LPC_HD bool pc_tree_ask_impl(const Acc& a, const int* words) {
  const int* p = words;
  return TreeForm<typename Acc::Univ, PC_TREE_FORM_DEPTH>::ask(a, p, false);
}
#ifdef LPC_HOST_HARNESS
template <class Acc> LPC_HD int pc_tree_deduce(Acc& a, const int* words) { return pc_tree_deduce_impl(a, words); }
template <class Acc> LPC_HD bool pc_tree_ask(const Acc& a, const int* words) { return pc_tree_ask_impl(a, words); }
#endif

} // namespace lpc
