// pc_oracle.cpp — TEST INFRASTRUCTURE ONLY (see pir_oracle.cpp). CPU restatement of the PC hot path of lala-pc:
// PC::deduce(i) -> pc::Formula::deduce -> pc::Term::{project, embed} over a VStore<Interval<ZLB>>.
//
// What it restates (paths relative to the lala-pc repository):
//   PC::deduce(int) / ask(int) / num_deductions       include/lala/pc.hpp:661-680
//   Formula as term: embed / project                  include/lala/formula.hpp:1090-1102
//   VariableLiteral<neg>                              include/lala/formula.hpp:80-167
//   Conjunction / Disjunction / Biconditional         include/lala/formula.hpp:241-451
//   Implication / ExclusiveDisjunction                include/lala/formula.hpp:453-587
//   Equality<neg> (=, !=)                             include/lala/formula.hpp:589-724
//   Inequality<neg> (<=, >)                           include/lala/formula.hpp:727-851
//   Constant, Variable, Unary<Neg|Abs>                include/lala/terms.hpp:18-175
//   Binary<GroupAdd|GroupSub|GroupMul<EDIV>>          include/lala/terms.hpp:177-262, 333-434
//   Binary<GroupMinMax<MIN|MAX>>                      include/lala/terms.hpp:301-331
//   Binary<GroupDiv<TDIV|FDIV|CDIV|EDIV>>, Nary<Mul>  include/lala/terms.hpp:264-299, 436-526
//   AbstractElement over the store                    include/lala/formula.hpp:14-77
//   Nary<Add>                                         include/lala/terms.hpp:436-526
// The leaf arithmetic, Interval::project(Sig, ...), additive_inverse, fjoin, LB::prev / UB::prev, lives in
// lala-core v1.2.8 (un-vendored, CMakeLists.txt:33-37) and is restated from its published behaviour:
//   ADD/SUB componentwise on the bounds (no bot check: Nary<Add>::embed feeds it the crossed interval produced by
//   additive_inverse, terms.hpp:190-194), NEG = [-ub,-lb], ABS, MUL / EDIV = hull of the corner results,
//   an infinite operand bound gives an infinite result bound.
//
// Parity status: PINNED by tests/pc_test.cpp for what its goldens exercise (TermTest.AddTermBinary/AddTermNary,
// <=, >, =, != on sums and differences, NEG, ABS, positive coefficients under <=, reification, clauses, infinite
// domains on a single variable). PARITY UNPINNED for: lower-bound rounding of `u ediv c` (reified sums with b = 0),
// negative coefficients, infinite bounds inside sums (tests/test_oracle_pc.py lists the pinned goldens).
//
// The walker is a template over the universe: Interval<ZLB> (Itv, tests/pc_test.cpp) and NBitset<64> (NBit,
// tests/pc_bitset_test.cpp:23-25). NBitset lives in lala-core too; restated: bit 0 = "some value <= -1", bit i (1..62) =
// value i-1, bit 63 = "some value >= 62" (pinned by `var -15..5` == NBit(-1,5), pc_bitset_test.cpp:152-156), meet = AND,
// join = OR, complement = NOT (`complemented`, formula.hpp:642-644), lb()/ub() = lowest / highest member with the end
// bits read as -inf / +inf, arithmetic (NEG, ABS, ADD, ...) through the interval [lb(), ub()] and back.
// Parity status of NBit: PINNED by the 10 goldens of pc_bitset_test.cpp (!=, =, clauses, int_abs);
// PARITY UNPINNED for NBitset arithmetic beyond IntAbs1 (the device path therefore only accepts the four pinned shapes
// on bitset stores, see include/lpc_pc.h).
//
// Formulas come in as a prefix-encoded int32 stream (see enum Tok); the same stream is what the tests flatten into
// the device encoding of include/lpc_pc.h.

#include <cstdint>
#include <climits>
#include <memory>
#include <vector>
#include <chrono>

namespace {

typedef int32_t v_t;
enum Tok { T_CONST = 1, T_VAR = 2, T_NEG = 3, T_ABS = 4, T_ADD = 5, T_SUB = 6, T_MUL = 7, T_NARY_ADD = 8, T_MIN = 9, T_MAX = 10,
           T_TDIV = 11, T_FDIV = 12, T_CDIV = 13, T_EDIV = 14, T_NARY_MUL = 15,
           F_VARLIT = 20, F_NVARLIT = 21, F_LEQ = 22, F_GT = 23, F_EQ = 24, F_NEQ = 25, F_AND = 26, F_OR = 27, F_EQUIV = 28,
           F_IMPLY = 29, F_XOR = 30, F_AE = 31, F_TRUE = 32, F_FALSE = 33 };   // True / False: formula.hpp:169-239
enum AeOp { AE_LEQ = 0, AE_GEQ = 1, AE_EQ = 2, AE_NEQ = 3 };   // F_AE op var k: a store-level element `var op k`
const v_t INF = INT32_MAX, MINF = INT32_MIN;

struct Itv {
  v_t lb, ub;
  static constexpr bool complemented = false;
  Itv() : lb(MINF), ub(INF) {}
  Itv(v_t l, v_t u) : lb(l), ub(u) {}
  bool is_bot() const { return lb > ub; }
  v_t lo() const { return lb; }
  v_t hi() const { return ub; }
  bool meet(const Itv& o) { bool c = false; if(o.lb > lb) { lb = o.lb; c = true; } if(o.ub < ub) { ub = o.ub; c = true; } return c; }
  bool contains0() const { return lb <= 0 && 0 <= ub; }          // `*this >= eq_zero` in the lattice order
  bool sub_of(v_t l, v_t u) const { return is_bot() || (lb >= l && ub <= u); }   // `*this <= [l,u]`
  bool same(const Itv& o) const { return (is_bot() && o.is_bot()) || (lb == o.lb && ub == o.ub); }
  Itv complement() const { return Itv(); }                        // never called (complemented == false)
  static Itv empty() { return Itv(INF, MINF); }
};
inline Itv fjoin(const Itv& a, const Itv& b) {
  if(a.is_bot()) return b;
  if(b.is_bot()) return a;
  return Itv(a.lb < b.lb ? a.lb : b.lb, a.ub > b.ub ? a.ub : b.ub);
}
inline bool inf(v_t x) { return x == INF || x == MINF; }
inline v_t clamp64(long long x) { return x >= INF ? INF : (x <= MINF ? MINF : (v_t)x); }
// bound arithmetic: an infinite operand gives an infinite result of the natural sign
inline v_t badd(v_t a, v_t b) { if(inf(a)) return a; if(inf(b)) return b; return clamp64((long long)a + b); }
inline v_t bneg(v_t a) { return a == INF ? MINF : (a == MINF ? INF : -a); }
inline v_t bsub(v_t a, v_t b) { return badd(a, bneg(b)); }
inline v_t bmul(v_t a, v_t b) {
  if(a == 0 || b == 0) return 0;
  if(inf(a) || inf(b)) return ((a < 0) != (b < 0)) ? MINF : INF;
  return clamp64((long long)a * b);
}
inline v_t bediv(v_t a, v_t b) {   // Euclidean division, b != 0
  if(inf(a)) return (b > 0) ? a : bneg(a);
  if(inf(b)) return 0;
  long long q = (long long)a / b, r = (long long)a % b;
  if(r < 0) q += (b > 0) ? -1 : 1;
  return clamp64(q);
}
inline Itv p_add(const Itv& a, const Itv& b) { return Itv(badd(a.lb, b.lb), badd(a.ub, b.ub)); }
inline Itv p_sub(const Itv& a, const Itv& b) { return Itv(bsub(a.lb, b.ub), bsub(a.ub, b.lb)); }
inline Itv p_neg(const Itv& a) { return Itv(bneg(a.ub), bneg(a.lb)); }
inline Itv additive_inverse(const Itv& a) { return Itv(bneg(a.lb), bneg(a.ub)); }
inline Itv p_abs(const Itv& a) {
  if(a.is_bot()) return a;
  if(a.lb >= 0) return a;
  if(a.ub <= 0) return p_neg(a);
  v_t m = bneg(a.lb) > a.ub ? bneg(a.lb) : a.ub;
  return Itv(0, m);
}
inline Itv hull4(v_t t1, v_t t2, v_t t3, v_t t4) {
  v_t lo = t1, hi = t1;
  for(v_t t : {t2, t3, t4}) { if(t < lo) lo = t; if(t > hi) hi = t; }
  return Itv(lo, hi);
}
inline Itv p_mul(const Itv& a, const Itv& b) {
  if(a.is_bot() || b.is_bot()) return Itv(INF, MINF);
  return hull4(bmul(a.lb, b.lb), bmul(a.lb, b.ub), bmul(a.ub, b.lb), bmul(a.ub, b.ub));
}
// EDIV by an interval that holds 0: the quotient is undefined there, so 0 is cut out of the divisor and the result is the
// hull over its negative and positive parts. Pinned for a 0 endpoint by IntTimes2 (pc_test.cpp:626-636: z = 1, y in [0,1]
// must give x = 1); a divisor straddling 0 or equal to {0} (-> empty) is UNPINNED.
inline Itv p_ediv_part(const Itv& a, v_t bl, v_t bu) {
  return hull4(bediv(a.lb, bl), bediv(a.lb, bu), bediv(a.ub, bl), bediv(a.ub, bu));
}
inline Itv p_ediv(const Itv& a, const Itv& b) {
  if(a.is_bot() || b.is_bot()) return Itv(INF, MINF);
  Itv r(INF, MINF);
  if(b.lb < 0) r = fjoin(r, p_ediv_part(a, b.lb, b.ub < -1 ? b.ub : -1));
  if(b.ub > 0) r = fjoin(r, p_ediv_part(a, b.lb > 1 ? b.lb : 1, b.ub));
  return r;
}

// project(TDIV | FDIV | CDIV | EDIV): the same shape with the rounding of the operator (0 cut out of the divisor).
// Pinned by IntDiv1-2 (pc_test.cpp:678-698) on 0/1 domains only; beyond that UNPINNED.
inline v_t bdivop(int kind, v_t a, v_t b) {   // b != 0
  if(kind == T_EDIV) return bediv(a, b);
  if(inf(a)) return (b > 0) ? a : bneg(a);
  if(inf(b)) return 0;
  long long q = (long long)a / b, r = (long long)a % b;
  if(kind == T_FDIV && r != 0 && ((r < 0) != (b < 0))) --q;
  if(kind == T_CDIV && r != 0 && ((r < 0) == (b < 0))) ++q;
  return clamp64(q);
}
inline Itv p_div(int kind, const Itv& a, const Itv& b) {
  if(a.is_bot() || b.is_bot()) return Itv(INF, MINF);
  Itv r(INF, MINF);
  if(b.lb < 0) { v_t bu = b.ub < -1 ? b.ub : -1; r = fjoin(r, hull4(bdivop(kind, a.lb, b.lb), bdivop(kind, a.lb, bu), bdivop(kind, a.ub, b.lb), bdivop(kind, a.ub, bu))); }
  if(b.ub > 0) { v_t bl = b.lb > 1 ? b.lb : 1; r = fjoin(r, hull4(bdivop(kind, a.lb, bl), bdivop(kind, a.lb, b.ub), bdivop(kind, a.ub, bl), bdivop(kind, a.ub, b.ub))); }
  return r;
}
// GroupDiv residuals (terms.hpp:272-296). left: x in u * y, upper bound joined with y.ub - 1 (join_ub(UB::prev(b.ub())));
// right: 0 leaves the ends of y, then y in x / u unless x holds 0 or u = {0}.
inline Itv div_left_residual(const Itv& u, const Itv& b) {
  Itv r = p_mul(u, b);
  if(r.is_bot()) return r;
  v_t pb = bsub(b.ub, 1);
  if(pb > r.ub) r.ub = pb;
  return r;
}
inline Itv div_right_residual(int kind, const Itv& u, const Itv& b, Itv r) {
  if(r.lb == 0) r.meet(Itv(1, INF));
  if(r.ub == 0) r.meet(Itv(MINF, -1));
  if(!b.contains0() && !(u.lb == 0 && u.ub == 0)) r.meet(p_div(kind, b, u));
  return r;
}

// project(MIN / MAX): componentwise on the bounds (pinned by MinConstraint1-3 / MaxConstraint1-3, pc_test.cpp:479-557)
inline Itv p_min(const Itv& a, const Itv& b) {
  if(a.is_bot() || b.is_bot()) return Itv(INF, MINF);
  return Itv(a.lb < b.lb ? a.lb : b.lb, a.ub < b.ub ? a.ub : b.ub);
}
inline Itv p_max(const Itv& a, const Itv& b) {
  if(a.is_bot() || b.is_bot()) return Itv(INF, MINF);
  return Itv(a.lb > b.lb ? a.lb : b.lb, a.ub > b.ub ? a.ub : b.ub);
}
inline Itv only_lb(const Itv& a) { return Itv(a.lb, INF); }
inline Itv only_ub(const Itv& a) { return Itv(MINF, a.ub); }

// NBitset<64, local_memory, unsigned long long> (lala-core, un-vendored; see the header).
struct NBit {
  uint64_t bits;
  static constexpr bool complemented = true;
  NBit() : bits(~0ull) {}
  NBit(v_t l, v_t u) : bits(0) {
    if(l > u) return;
    const int from = l < 0 ? 0 : (l >= 62 ? 63 : l + 1), to = u < 0 ? 0 : (u >= 62 ? 63 : u + 1);
    bits = (~0ull << from) & (~0ull >> (63 - to));
  }
  static NBit raw(uint64_t b) { NBit r; r.bits = b; return r; }
  bool is_bot() const { return bits == 0; }
  v_t lo() const { return (bits & 1) ? MINF : (bits ? (v_t)__builtin_ctzll(bits) - 1 : INF); }
  v_t hi() const { return (bits >> 63) ? INF : (bits ? 62 - (v_t)__builtin_clzll(bits) : MINF); }
  bool meet(const NBit& o) { const uint64_t n = bits & o.bits; const bool c = n != bits; bits = n; return c; }
  bool contains0() const { return (bits >> 1) & 1; }              // superset of eq_zero = {0}
  bool sub_of(v_t l, v_t u) const { return (bits & ~NBit(l, u).bits) == 0; }
  bool same(const NBit& o) const { return bits == o.bits; }
  NBit complement() const { return raw(~bits); }
  static NBit empty() { return raw(0); }
  Itv itv() const { return Itv(lo(), hi()); }
};
inline NBit from_itv(const Itv& i) { return NBit(i.lb, i.ub); }
inline NBit fjoin(const NBit& a, const NBit& b) { return NBit::raw(a.bits | b.bits); }
inline NBit p_add(const NBit& a, const NBit& b) { return (a.is_bot() || b.is_bot()) ? NBit::raw(0) : from_itv(p_add(a.itv(), b.itv())); }
inline NBit p_sub(const NBit& a, const NBit& b) { return (a.is_bot() || b.is_bot()) ? NBit::raw(0) : from_itv(p_sub(a.itv(), b.itv())); }
inline NBit p_neg(const NBit& a) { return a.is_bot() ? a : from_itv(p_neg(a.itv())); }
inline NBit additive_inverse(const NBit& a) { return p_neg(a); }   // unpinned; a set has no crossed form
inline NBit p_abs(const NBit& a) { return a.is_bot() ? a : from_itv(p_abs(a.itv())); }
inline NBit p_mul(const NBit& a, const NBit& b) { return (a.is_bot() || b.is_bot()) ? NBit::raw(0) : from_itv(p_mul(a.itv(), b.itv())); }
inline NBit p_min(const NBit& a, const NBit& b) { return (a.is_bot() || b.is_bot()) ? NBit::raw(0) : from_itv(p_min(a.itv(), b.itv())); }
inline NBit p_max(const NBit& a, const NBit& b) { return (a.is_bot() || b.is_bot()) ? NBit::raw(0) : from_itv(p_max(a.itv(), b.itv())); }
inline NBit only_lb(const NBit& a) { return from_itv(Itv(a.lo(), INF)); }   // unpinned on bitsets
inline NBit only_ub(const NBit& a) { return from_itv(Itv(MINF, a.hi())); }
inline NBit p_div(int kind, const NBit& a, const NBit& b) { return (a.is_bot() || b.is_bot()) ? NBit::raw(0) : from_itv(p_div(kind, a.itv(), b.itv())); }
inline NBit div_left_residual(const NBit& u, const NBit& b) { return (u.is_bot() || b.is_bot()) ? NBit::raw(0) : from_itv(div_left_residual(u.itv(), b.itv())); }
inline NBit div_right_residual(int kind, const NBit& u, const NBit& b, NBit r) { return r.is_bot() ? r : from_itv(div_right_residual(kind, u.itv(), b.itv(), r.itv())); }
inline NBit p_ediv(const NBit& a, const NBit& b) { return (a.is_bot() || b.is_bot()) ? NBit::raw(0) : from_itv(p_ediv(a.itv(), b.itv())); }

template <class U>
struct Store {
  U* d; int n; bool bot;     // 8-byte cells: {int32 lb, int32 ub} or one uint64 bitset
  bool touched = false;      // some embed changed a cell since this was last cleared (see fixpoint())
  const U& get(int v) const { return d[v]; }
  void project(int v, U& r) const { r.meet(d[v]); }
  bool embed(int v, const U& u) {                                    // VStore::embed
    if(d[v].meet(u)) { touched = true; if(d[v].is_bot()) bot = true; return true; }
    return false;
  }
  bool is_bot() const { return bot; }
  // a.meet_bot() (formula.hpp:185-193): the whole element fails. VStore keeps a flag; the device has no flag and empties
  // variable 0 - stores of failed elements are not compared (only that they failed).
  bool meet_bot() {
    if(bot) return false;
    bot = true; touched = true;
    if(n > 0) d[0] = U::empty();
    return true;
  }
};
static_assert(sizeof(Itv) == 8 && sizeof(NBit) == 8, "8-byte cells");



template <class U>
struct Term {
  int kind = 0; v_t k = 0; int var = -1;
  std::vector<std::unique_ptr<Term>> sub;
  bool is_const() const { return kind == T_CONST; }

  void project(const Store<U>& a, U& r) const {
    switch(kind) {
      case T_CONST: r.meet(U(k, k)); break;                          // terms.hpp:34
      case T_VAR: a.project(var, r); break;                          // terms.hpp:69-71
      case T_NEG: { U t; sub[0]->project(a, t); r.meet(p_neg(t)); break; }   // terms.hpp:148-152, 91-93
      case T_ABS: { U t; sub[0]->project(a, t); r.meet(p_abs(t)); break; }
      case T_ADD: case T_SUB: case T_MUL: case T_MIN: case T_MAX: {   // terms.hpp:399-405
        U x, y; sub[0]->project(a, x); sub[1]->project(a, y);
        r.meet(kind == T_ADD ? p_add(x, y) : kind == T_SUB ? p_sub(x, y) : kind == T_MUL ? p_mul(x, y)
               : kind == T_MIN ? p_min(x, y) : p_max(x, y));
        break;
      }
      case T_TDIV: case T_FDIV: case T_CDIV: case T_EDIV: {           // GroupDiv::project, terms.hpp:268-270
        U x, y; sub[0]->project(a, x); sub[1]->project(a, y);
        r.meet(p_div(kind, x, y));
        break;
      }
      case T_NARY_MUL: {                                             // Nary<Mul>, terms.hpp:465-478
        U accu; sub[0]->project(a, accu);
        for(size_t i = 1; i < sub.size(); ++i) { U t; sub[i]->project(a, t); accu = p_mul(accu, t); }
        r.meet(accu);
        break;
      }
      case T_NARY_ADD: {                                             // terms.hpp:465-478
        U accu; sub[0]->project(a, accu);
        for(size_t i = 1; i < sub.size(); ++i) { U t; sub[i]->project(a, t); accu = p_add(accu, t); }
        r.meet(accu);
        break;
      }
    }
  }

  bool embed(Store<U>& a, const U& u) const {
    switch(kind) {
      case T_CONST: return false;                                    // terms.hpp:33
      case T_VAR: return a.embed(var, u);                            // terms.hpp:65-67
      case T_NEG: { U t; t.meet(p_neg(u)); return sub[0]->embed(a, t); }     // terms.hpp:142-146, 95-97
      case T_ABS: { U t; t.meet(fjoin(u, p_neg(u))); return sub[0]->embed(a, t); }   // terms.hpp:112-114
      case T_ADD: case T_SUB: case T_MUL: case T_MIN: case T_MAX: {   // terms.hpp:376-397
        bool ch = false;
        if(!sub[0]->is_const()) {
          U yt, res; sub[1]->project(a, yt);
          left_residual(u, yt, res);
          ch |= sub[0]->embed(a, res);
        }
        if(!sub[1]->is_const()) {
          U xt, res; sub[0]->project(a, xt);
          right_residual(u, xt, res);
          ch |= sub[1]->embed(a, res);
        }
        return ch;
      }
      case T_TDIV: case T_FDIV: case T_CDIV: case T_EDIV: {           // Binary::embed with a division group, terms.hpp:376-397
        bool ch = false;
        if(!sub[0]->is_const()) {
          U yt; sub[1]->project(a, yt);
          U res; res.meet(div_left_residual(u, yt));
          ch |= sub[0]->embed(a, res);
        }
        if(!sub[1]->is_const()) {
          U xt; sub[0]->project(a, xt);
          U cur; sub[1]->project(a, cur);                            // "we read it to potentially remove 0" (:389-392)
          ch |= sub[1]->embed(a, div_right_residual(kind, u, xt, cur));
        }
        return ch;
      }
      case T_NARY_MUL: {                                             // Nary<Mul>::embed, terms.hpp:480-499
        U all; project(a, all);
        bool ch = false;
        if(all.same(U(0, 0))) return false;                          // GroupMul::is_absorbing (:239-241)
        for(size_t i = 0; i < sub.size(); ++i) {
          U tmp; sub[i]->project(a, tmp);
          U tmp2; tmp2.meet(p_ediv(all, tmp));                       // rev_op (:244-246)
          U res;
          if(!(u.contains0() && tmp2.contains0())) res.meet(p_ediv(u, tmp2));   // left_residual (:249-253)
          ch |= sub[i]->embed(a, res);
        }
        return ch;
      }
      case T_NARY_ADD: {                                             // terms.hpp:480-499
        U all; project(a, all);
        bool ch = false;
        for(size_t i = 0; i < sub.size(); ++i) {
          U tmp; sub[i]->project(a, tmp);
          U tmp2; tmp2.meet(p_add(all, additive_inverse(tmp)));      // GroupAdd::rev_op, terms.hpp:190-194
          U res; res.meet(p_sub(u, tmp2));                           // GroupAdd::left_residual, :196-198
          ch |= sub[i]->embed(a, res);
        }
        return ch;
      }
    }
    return false;
  }

  void left_residual(const U& u, const U& b, U& r) const {
    if(kind == T_ADD) r.meet(p_sub(u, b));                           // terms.hpp:196-198
    else if(kind == T_SUB) r.meet(p_add(u, b));                      // terms.hpp:218-220
    else if(kind == T_MIN || kind == T_MAX) {                        // GroupMinMax, terms.hpp:310-322
      U m = u; m.meet(b);
      if(m.is_bot()) r.meet(u);                                      // the other operand cannot be the result
      else r.meet(kind == T_MIN ? only_lb(u) : only_ub(u));
    }
    else if(!(u.contains0() && b.contains0())) r.meet(p_ediv(u, b)); // GroupMul, terms.hpp:249-253
  }
  void right_residual(const U& u, const U& b, U& r) const {
    if(kind == T_SUB) r.meet(p_sub(b, u));                           // terms.hpp:222-224
    else left_residual(u, b, r);
  }
};

template <class U>
struct Formula {
  int kind = 0; int var = -1;
  int ae_op = 0; v_t ae_k = 0;     // F_AE
  std::unique_ptr<Term<U>> l, r;
  std::unique_ptr<Formula> f, g;

  bool ask(const Store<U>& a) const { return ask_impl(a, false); }
  bool nask(const Store<U>& a) const { return ask_impl(a, true); }
  bool deduce(Store<U>& a) const { return deduce_impl(a, false); }
  bool contradeduce(Store<U>& a) const { return deduce_impl(a, true); }

  // `negated` selects the dual operation (nask / contradeduce); literal and comparison kinds fold it into `neg`.
  bool ask_impl(const Store<U>& a, bool negated) const {
    switch(kind) {
      case F_VARLIT: case F_NVARLIT: {                               // formula.hpp:97-110, 126-135
        bool neg = (kind == F_NVARLIT) != negated;
        U t; a.project(var, t);
        return neg ? t.sub_of(0, 0) : !t.contains0();
      }
      case F_LEQ: case F_GT: {                                       // formula.hpp:757-771
        bool neg = (kind == F_GT) != negated;
        U x, y; l->project(a, x); r->project(a, y);
        return neg ? x.lo() > y.hi() : x.hi() <= y.lo();
      }
      case F_EQ: case F_NEQ: {                                       // formula.hpp:616-631
        bool neg = (kind == F_NEQ) != negated;
        U x, y; l->project(a, x); r->project(a, y);
        if(neg) { U m = x; m.meet(y); return m.is_bot(); }
        return x.same(y) && x.lo() == x.hi();
      }
      case F_TRUE: return negated ? a.is_bot() : true;              // formula.hpp:220-221
      case F_FALSE: return negated ? true : a.is_bot();             // formula.hpp:182-183
      case F_AE: {   // AbstractElement::ask / nask = the store's ask of the element / of its negation (formula.hpp:43-49)
        int op = ae_op; v_t k = ae_k;
        if(negated) { if(op == AE_LEQ) { op = AE_GEQ; k = badd(k, 1); } else if(op == AE_GEQ) { op = AE_LEQ; k = bsub(k, 1); } else op = op == AE_EQ ? AE_NEQ : AE_EQ; }
        const U& d = a.get(var);
        switch(op) {
          case AE_LEQ: return d.sub_of(MINF, k);
          case AE_GEQ: return d.sub_of(k, INF);
          case AE_EQ: return d.sub_of(k, k);
          default: { U m = d; m.meet(U(k, k)); return m.is_bot(); }
        }
      }
      case F_AND: return negated ? (f->nask(a) || g->nask(a)) : (f->ask(a) && g->ask(a));       // formula.hpp:268-274
      case F_OR: return negated ? (f->nask(a) && g->nask(a)) : (f->ask(a) || g->ask(a));        // formula.hpp:338-344
      case F_EQUIV:                                                  // formula.hpp:408-419
        return negated ? ((f->ask(a) && g->nask(a)) || (f->nask(a) && g->ask(a)))
                       : ((f->ask(a) && g->ask(a)) || (f->nask(a) && g->nask(a)));
      case F_IMPLY:                                                  // formula.hpp:479-487
        return negated ? (f->ask(a) && g->nask(a)) : (f->nask(a) || g->ask(a));
      case F_XOR:                                                    // formula.hpp:541-552 (nask exactly as written there)
        return negated ? ((f->ask(a) && g->ask(a)) || (f->ask(a) && g->nask(a)))
                       : ((f->ask(a) && g->nask(a)) || (f->nask(a) && g->ask(a)));
    }
    return false;
  }

  // Equality<true>::deduce for one direction: `other` loses the singleton `single` (formula.hpp:640-653, 656-669)
  static bool shave(Store<U>& a, const Term<U>& other, const U& single) {
    if(U::complemented) return other.embed(a, single.complement());
    U o; other.project(a, o);
    U lo = o, hi = o;
    lo.meet(U(badd(single.lo(), 1), INF));
    hi.meet(U(MINF, bsub(single.hi(), 1)));
    return other.embed(a, fjoin(lo, hi));
  }

  bool deduce_impl(Store<U>& a, bool negated) const {
    switch(kind) {
      case F_VARLIT: case F_NVARLIT: {                               // formula.hpp:112-120, 140-149
        bool neg = (kind == F_NVARLIT) != negated;
        return a.embed(var, neg ? U(0, 0) : U(1, 1));
      }
      case F_LEQ: case F_GT: {                                       // formula.hpp:773-807
        bool neg = (kind == F_GT) != negated;
        bool ch = false;
        U x, y;
        if(neg) {   // l > r
          if(!l->is_const()) { r->project(a, y); y.meet(U(badd(y.lo(), 1), INF)); ch = l->embed(a, U(y.lo(), INF)); }
          if(!r->is_const()) { l->project(a, x); x.meet(U(MINF, bsub(x.hi(), 1))); ch |= r->embed(a, U(MINF, x.hi())); }
        }
        else {      // l <= r
          if(!l->is_const()) { r->project(a, y); ch |= l->embed(a, U(MINF, y.hi())); }
          if(!r->is_const()) { l->project(a, x); ch = r->embed(a, U(x.lo(), INF)); }   // `=` as in formula.hpp:803
        }
        return ch;
      }
      case F_EQ: case F_NEQ: {                                       // formula.hpp:633-683
        bool neg = (kind == F_NEQ) != negated;
        U x, y;
        if(neg) {
          if(!r->is_const()) {
            l->project(a, x);
            if(x.lo() == x.hi()) return shave(a, *r, x);
          }
          if(!l->is_const()) {
            r->project(a, y);
            if(y.lo() == y.hi()) return shave(a, *l, y);
          }
          return false;
        }
        bool ch = false;
        if(!r->is_const()) { l->project(a, x); ch = r->embed(a, x); }
        if(!l->is_const()) { r->project(a, y); ch |= l->embed(a, y); }
        return ch;
      }
      case F_TRUE: return negated ? a.meet_bot() : false;           // formula.hpp:222-228
      case F_FALSE: return negated ? false : a.meet_bot();          // formula.hpp:185-193
      case F_AE: {   // AbstractElement::deduce / contradeduce = the store's deduce of the element / its negation
        int op = ae_op; v_t k = ae_k;   // (formula.hpp:51-57); `!=` has no interval (AbstractElement3-4, pc_test.cpp:738-764)
        if(negated) { if(op == AE_LEQ) { op = AE_GEQ; k = badd(k, 1); } else if(op == AE_GEQ) { op = AE_LEQ; k = bsub(k, 1); } else op = op == AE_EQ ? AE_NEQ : AE_EQ; }
        switch(op) {
          case AE_LEQ: return a.embed(var, U(MINF, k));
          case AE_GEQ: return a.embed(var, U(k, INF));
          case AE_EQ: return a.embed(var, U(k, k));
          default: return U::complemented ? a.embed(var, U(k, k).complement()) : false;
        }
      }
      case F_AND:                                                    // formula.hpp:276-286
        if(!negated) { bool c = f->deduce(a); c |= g->deduce(a); return c; }
        if(f->ask(a)) return g->contradeduce(a);
        else if(g->ask(a)) return f->contradeduce(a);
        return false;
      case F_OR:                                                     // formula.hpp:346-356
        if(negated) { bool c = f->contradeduce(a); c |= g->contradeduce(a); return c; }
        if(f->nask(a)) return g->deduce(a);
        else if(g->nask(a)) return f->deduce(a);
        return false;
      case F_EQUIV:                                                  // formula.hpp:421-435
        if(!negated) {
          if(f->ask(a)) return g->deduce(a);
          else if(f->nask(a)) return g->contradeduce(a);
          else if(g->ask(a)) return f->deduce(a);
          else if(g->nask(a)) return f->contradeduce(a);
          return false;
        }
        if(f->ask(a)) return g->contradeduce(a);
        else if(f->nask(a)) return g->deduce(a);
        else if(g->ask(a)) return f->contradeduce(a);
        else if(g->nask(a)) return f->deduce(a);
        return false;
      case F_IMPLY:                                                  // formula.hpp:489-499
        if(f->ask(a)) return negated ? g->contradeduce(a) : g->deduce(a);
        else if(g->nask(a)) return negated ? f->deduce(a) : f->contradeduce(a);
        return false;
      case F_XOR:                                                    // formula.hpp:554-568
        if(f->ask(a)) return negated ? g->deduce(a) : g->contradeduce(a);
        else if(f->nask(a)) return negated ? g->contradeduce(a) : g->deduce(a);
        else if(g->ask(a)) return negated ? f->deduce(a) : f->contradeduce(a);
        else if(g->nask(a)) return negated ? f->contradeduce(a) : f->deduce(a);
        return false;
    }
    return false;
  }
};

template <class U>
std::unique_ptr<Term<U>> parse_term(const int32_t*& p) {
  std::unique_ptr<Term<U>> t(new Term<U>());
  t->kind = *p++;
  switch(t->kind) {
    case T_CONST: t->k = *p++; break;
    case T_VAR: t->var = *p++; break;
    case T_NEG: case T_ABS: t->sub.push_back(parse_term<U>(p)); break;
    case T_ADD: case T_SUB: case T_MUL: case T_MIN: case T_MAX: case T_TDIV: case T_FDIV: case T_CDIV: case T_EDIV:
      t->sub.push_back(parse_term<U>(p)); t->sub.push_back(parse_term<U>(p)); break;
    case T_NARY_ADD: case T_NARY_MUL: { int n = *p++; for(int i = 0; i < n; ++i) t->sub.push_back(parse_term<U>(p)); break; }
  }
  return t;
}
template <class U>
std::unique_ptr<Formula<U>> parse_formula(const int32_t*& p) {
  std::unique_ptr<Formula<U>> f(new Formula<U>());
  f->kind = *p++;
  switch(f->kind) {
    case F_VARLIT: case F_NVARLIT: f->var = *p++; break;
    case F_AE: f->ae_op = *p++; f->var = *p++; f->ae_k = *p++; break;
    case F_TRUE: case F_FALSE: break;
    case F_LEQ: case F_GT: case F_EQ: case F_NEQ: f->l = parse_term<U>(p); f->r = parse_term<U>(p); break;
    default: f->f = parse_formula<U>(p); f->g = parse_formula<U>(p); break;
  }
  return f;
}

// One model keeps both instantiations of its propagators; `bits` selects the universe of the store it is run on.
struct Model {
  std::vector<std::unique_ptr<Formula<Itv>>> props;
  std::vector<std::unique_ptr<Formula<NBit>>> bprops;
};

template <class U> bool scan_bot(const U* d, int n) { for(int i = 0; i < n; ++i) if(d[i].is_bot()) return true; return false; }

struct lpco_stats_ { int32_t has_changed, is_bot; int64_t sweeps, deductions; double seconds; };

// GaussSeidelIteration::fixpoint over PC::deduce(i) (tests/pc_test.cpp:91-94, pc_bitset_test.cpp:52-55).
template <class U>
void fixpoint(const std::vector<std::unique_ptr<Formula<U>>>& props, U* cells, int nvars, int stop_on_bot,
              int64_t max_sweeps, lpco_stats_* out) {
  Store<U> s{cells, nvars, scan_bot(cells, nvars)};
  auto t0 = std::chrono::steady_clock::now();
  bool changed = true, any = false;
  int64_t sweeps = 0;
  const size_t n = props.size();
  while(changed && !(stop_on_bot && s.bot) && (max_sweeps <= 0 || sweeps < max_sweeps)) {
    changed = false;
    s.touched = false;
    for(size_t i = 0; i < n; ++i) changed |= props[i]->deduce(s);
    // Inequality::deduce ASSIGNS its flag on the right-hand side (formula.hpp:803), so `l <= r` with two non-constant
    // sides can tighten l and return false; a loop driven by the flags alone may then stop short of a fixpoint. The
    // checker keeps iterating while the store moves - identical on every reference golden (none hits the case), and the
    // only well-defined target for a parallel schedule. deduce(i)'s own return value stays literal.
    changed |= s.touched;
    any |= changed;
    ++sweeps;
  }
  auto t1 = std::chrono::steady_clock::now();
  if(out) { out->has_changed = any; out->is_bot = s.bot; out->sweeps = sweeps; out->deductions = sweeps * (int64_t)n;
            out->seconds = std::chrono::duration<double>(t1 - t0).count(); }
}

template <class U>
int64_t ask_all(const std::vector<std::unique_ptr<Formula<U>>>& props, const U* cells, int nvars, uint8_t* bits) {
  Store<U> s{const_cast<U*>(cells), nvars, false};
  int64_t c = 0;
  for(size_t i = 0; i < props.size(); ++i) { bool e = props[i]->ask(s); if(bits) bits[i] = e; c += e; }
  return c;
}

} // namespace

extern "C" {

typedef lpco_stats_ lpco_stats;

// stream: n_props formulas, prefix encoded back to back. Returns an opaque model.
void* lpco_pc_parse(const int32_t* stream, int32_t n_props) {
  Model* m = new Model();
  const int32_t* p = stream;
  for(int i = 0; i < n_props; ++i) m->props.push_back(parse_formula<Itv>(p));
  p = stream;
  for(int i = 0; i < n_props; ++i) m->bprops.push_back(parse_formula<NBit>(p));
  return m;
}
void lpco_pc_free(void* m) { delete static_cast<Model*>(m); }

// PC::deduce(i) (pc.hpp:671-680) on an interleaved {lb,ub} store.
int lpco_pc_deduce(void* m, int32_t i, int32_t* lbub, int32_t nvars, int32_t* is_bot) {
  Store<Itv> s{reinterpret_cast<Itv*>(lbub), nvars, is_bot && *is_bot};
  bool c = static_cast<Model*>(m)->props[i]->deduce(s);
  if(is_bot) *is_bot = s.bot;
  return c;
}
int lpco_pc_ask(void* m, int32_t i, const int32_t* lbub, int32_t nvars) {
  Store<Itv> s{reinterpret_cast<Itv*>(const_cast<int32_t*>(lbub)), nvars, false};
  return static_cast<Model*>(m)->props[i]->ask(s);
}
// The same two on a VStore<NBitset<64>> (one uint64 per variable).
int lpco_pc_deduce_bits(void* m, int32_t i, uint64_t* cells, int32_t nvars, int32_t* is_bot) {
  Store<NBit> s{reinterpret_cast<NBit*>(cells), nvars, is_bot && *is_bot};
  bool c = static_cast<Model*>(m)->bprops[i]->deduce(s);
  if(is_bot) *is_bot = s.bot;
  return c;
}
int lpco_pc_ask_bits(void* m, int32_t i, const uint64_t* cells, int32_t nvars) {
  Store<NBit> s{reinterpret_cast<NBit*>(const_cast<uint64_t*>(cells)), nvars, false};
  return static_cast<Model*>(m)->bprops[i]->ask(s);
}
// Term-level entry points for TermTest.* (pc_test.cpp:31-67): project / embed of ONE term stream.
void lpco_pc_term_project(const int32_t* stream, const int32_t* lbub, int32_t nvars, int32_t* out2) {
  const int32_t* p = stream;
  auto t = parse_term<Itv>(p);
  Store<Itv> s{reinterpret_cast<Itv*>(const_cast<int32_t*>(lbub)), nvars, false};
  Itv r; t->project(s, r);
  out2[0] = r.lb; out2[1] = r.ub;
}
int lpco_pc_term_embed(const int32_t* stream, int32_t* lbub, int32_t nvars, int32_t lb, int32_t ub) {
  const int32_t* p = stream;
  auto t = parse_term<Itv>(p);
  Store<Itv> s{reinterpret_cast<Itv*>(lbub), nvars, false};
  return t->embed(s, Itv(lb, ub));
}

// GaussSeidelIteration::fixpoint over PC::deduce(i); stop_on_bot as in pir_oracle.cpp.
void lpco_pc_fixpoint(void* mp, int32_t* lbub, int32_t nvars, int32_t stop_on_bot, int64_t max_sweeps, lpco_stats* out) {
  fixpoint(static_cast<Model*>(mp)->props, reinterpret_cast<Itv*>(lbub), nvars, stop_on_bot, max_sweeps, out);
}
void lpco_pc_fixpoint_bits(void* mp, uint64_t* cells, int32_t nvars, int32_t stop_on_bot, int64_t max_sweeps, lpco_stats* out) {
  fixpoint(static_cast<Model*>(mp)->bprops, reinterpret_cast<NBit*>(cells), nvars, stop_on_bot, max_sweeps, out);
}

int64_t lpco_pc_ask_all(void* mp, const int32_t* lbub, int32_t nvars, uint8_t* bits) {
  return ask_all(static_cast<Model*>(mp)->props, reinterpret_cast<const Itv*>(lbub), nvars, bits);
}
int64_t lpco_pc_ask_all_bits(void* mp, const uint64_t* cells, int32_t nvars, uint8_t* bits) {
  return ask_all(static_cast<Model*>(mp)->bprops, reinterpret_cast<const NBit*>(cells), nvars, bits);
}

// NBitset(lb, ub) constructor and lb()/ub() for the tests' conversions (pc_bitset_test.cpp: NBit(1,9), from_set).
uint64_t lpco_nbit(int32_t lb, int32_t ub) { return NBit(lb, ub).bits; }
void lpco_nbit_bounds(uint64_t bits, int32_t* out2) { NBit b = NBit::raw(bits); out2[0] = b.lo(); out2[1] = b.hi(); }

} // extern "C"
