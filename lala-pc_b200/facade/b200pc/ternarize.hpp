// b200pc/ternarize.hpp — host front-end of the PIR path: rewrite integer constraints into the ternary form
// `X = Y op Z` that PIR::interpret_tell reads (lala-pc include/lala/pir.hpp:254-287).
//
// The reference gets there through lala-core's `ternarize` (lala/logic/ternarize.hpp, included at pir.hpp:15,
// un-vendored). This header restates that STEP, not its text: the decomposition is the textbook one and the names of the
// variables it introduces are its own, so what the tests check is semantic (the reference's expected intervals for the
// FlatZinc models of tests/pir_test.cpp, which go through upstream's ternariser there and through this one here).
//   * a constant becomes a singleton variable (`ONE`, `ZERO`, `C5` in pir_test.cpp's hand-written networks);
//   * every non-variable sub-term gets a fresh variable `t = a op b`, identical definitions share one variable;
//     `a - b` is `a = d + b`, `-a` is `0 = n + a`, `|a|` is `max(a, -a)`, an n-ary sum is a left-nested chain;
//   * comparisons and connectives are reified into 0/1 variables: `r = (a <= b)`, `r = (a == b)`, and = min, or = max,
//     not r = 1 - r (`1 = r + s`), imply = (r1 <= r2), equiv = (r1 == r2), xor = not equiv;
//   * a constraint asserted at top level uses the constant ONE (ZERO for its negation) as its result variable, which is
//     how the reference's tests write `x <= y`: LEQ(X = ONE, Y = x, Z = y);
//   * a temporary starts at the interval hull of its definition, not at top: a top temporary meets rules such as
//     `x = 0 => y >= z.lb + 1` (pir.hpp:750-753) that turn an infinite bound into a finite one next to INT_MIN, after
//     which `+` wraps (the reference is undefined there) and the result depends on the evaluation order.
// Input: the TF trees of pc.hpp. Output: ternary formulas in pir.hpp's F type plus the domains of the new variables,
// ready for PIR::interpret_tell. Python counterpart: lala-pc_b200/ternarize.py (same decomposition).
#pragma once
#include <map>
#include <tuple>

#include "pc.hpp"

namespace b200pc {

class Ternarizer {
public:
  // `env` receives the new variables; `dom(name)` must give the domain of every variable of the input formulas.
  Ternarizer(VarEnv& env, std::map<std::string, Itv> doms) : env_(env), doms_(std::move(doms)) {}

  // Assert that `f` holds (or, with truth = false, that it does not). Returns false on a node without a ternary form.
  // Top-level shapes are written the way the reference's ternarised models come out (deduction counts of
  // tests/pir_test.cpp): a comparison with a constant is a bound on the term's variable (`int_le(int_plus(x, y), 5)` is
  // ONE propagator, t = x + y, with t <= 5 told to the store), `b = (x <= 5)` is one record, a clause ends in
  // `ONE = max(.., ..)`, xor / equiv / imply compare the two reified sides directly.
  bool tell(const TF& f, bool truth = true) {
    if(f.is_variable()) return bound(f.name, EQ, truth ? 1 : 0);
    if(!f.is_seq()) return false;
    const int sig = f.sig();
    if(sig == NOT && f.args.size() == 1) return tell(f.seq(0), !truth);
    if(sig == AND && truth) { for(auto& a : f.args) if(!tell(a, true)) return false; return true; }
    if(sig == OR && !truth) { for(auto& a : f.args) if(!tell(a, false)) return false; return true; }
    if((sig == OR || sig == AND) && f.args.size() >= 2) {   // a clause: ONE = max(max(l1, l2), l3) (ZERO = min(..) for a refuted and)
      std::string acc, b;
      if(!reify(f.seq(0), acc)) return false;
      for(size_t i = 1; i + 1 < f.args.size(); ++i) { if(!reify(f.seq((int)i), b)) return false; acc = define(sig == OR ? MAX : MIN, acc, b, true); }
      if(!reify(f.args.back(), b)) return false;
      return emit(sig == OR ? MAX : MIN, konst(sig == OR ? 1 : 0), acc, b);
    }
    if(!f.is_binary()) return false;
    const TF& a = f.seq(0); const TF& b = f.seq(1);
    const bool logical = a.is_logical() || a.is_predicate() || b.is_logical() || b.is_predicate();
    std::string x, y;
    int cmp = sig;
    if(cmp == EQ && logical) cmp = EQUIV;
    switch(cmp) {
      case LEQ: case GEQ: case GT: case LT: case EQ: case NEQ: {
        if(!truth) cmp = cmp == LEQ ? GT : cmp == GT ? LEQ : cmp == GEQ ? LT : cmp == LT ? GEQ : cmp == EQ ? NEQ : EQ;
        // a bound on one variable goes to the store
        if(b.is_constant() && !a.is_constant() && cmp != NEQ) return term(a, x) && bound(x, cmp, b.k);
        if(a.is_constant() && !b.is_constant() && cmp != NEQ)
          return term(b, y) && bound(y, cmp == LEQ ? GEQ : cmp == GEQ ? LEQ : cmp == GT ? LT : cmp == LT ? GT : EQ, a.k);
        if(cmp == EQ) {   // x = y op z is already ternary (either orientation)
          const TF* v = a.is_variable() ? &a : b.is_variable() ? &b : nullptr;
          const TF* t = v == &a ? &b : &a;
          if(v && t->is_binary() && pir_op(t->sig()) && t->sig() != EQ && t->sig() != LEQ)
            return term(t->seq(0), x) && term(t->seq(1), y) && emit(t->sig(), v->name, x, y);
          if(v && t->is_seq() && t->sig() == ABS && t->args.size() == 1)   // y = |a|: y = max(a, -a) and y >= 0 (IntAbs1: 3 propagators)
            return term(t->seq(0), x) && emit(MAX, v->name, x, define(SUB, konst(0), x, false)) && emit(LEQ, konst(1), konst(0), v->name);
        }
        if(!term(a, x) || !term(b, y)) return false;
        switch(cmp) {
          case LEQ: return emit(LEQ, konst(1), x, y);
          case GEQ: return emit(LEQ, konst(1), y, x);
          case GT: return emit(LEQ, konst(0), x, y);
          case LT: return emit(LEQ, konst(0), y, x);
          case EQ: return emit(EQ, konst(1), x, y);
          default: return emit(EQ, konst(0), x, y);
        }
      }
      case EQUIV: case XOR: {
        const bool same = (cmp == EQUIV) == truth;
        if(same) {   // b <=> (y <= z) / (y = z) with b a variable is one reified record
          const TF* v = a.is_variable() ? &a : b.is_variable() ? &b : nullptr;
          const TF* t = v == &a ? &b : &a;
          if(v && doms_.count(v->name) && t->is_binary() && !(t->seq(0).is_logical() || t->seq(0).is_predicate() || t->seq(1).is_logical() || t->seq(1).is_predicate())) {
            if(t->sig() == LEQ) return term(t->seq(0), x) && term(t->seq(1), y) && emit(LEQ, v->name, x, y);
            if(t->sig() == GEQ) return term(t->seq(0), x) && term(t->seq(1), y) && emit(LEQ, v->name, y, x);
            if(t->sig() == EQ) return term(t->seq(0), x) && term(t->seq(1), y) && emit(EQ, v->name, x, y);
          }
          if(v && doms_.count(v->name) && t->is_seq() && (t->sig() == AND || t->sig() == OR) && t->args.size() >= 2) {
            std::string acc, last;   // b <=> (p /\ q): b = min(rp, rq) (ResourceConstraint1: 5 propagators)
            if(!reify(t->seq(0), acc)) return false;
            for(size_t i = 1; i + 1 < t->args.size(); ++i) { if(!reify(t->seq((int)i), last)) return false; acc = define(t->sig() == AND ? MIN : MAX, acc, last, true); }
            return reify(t->args.back(), last) && emit(t->sig() == AND ? MIN : MAX, v->name, acc, last);
          }
        }
        return reify(a, x) && reify(b, y) && emit(EQ, konst(same ? 1 : 0), x, y);
      }
      case IMPLY:
        if(truth) return reify(a, x) && reify(b, y) && emit(LEQ, konst(1), x, y);
        return tell(a, true) && tell(b, false);   // not (a => b) == a and not b
      default: break;
    }
    std::string r;
    if(!reify(f, r)) return false;
    return bound(r, EQ, truth ? 1 : 0);
  }

  // The ternary constraints (`X = Y op Z` as EQ(var, binary(var, op, var))) and the unary bounds (`x <op> k`), in emission
  // order; each is accepted by PIR::interpret_tell.
  const std::vector<F>& constraints() const { return out_; }
  // The variables this object declared, with their initial domains (tell them to the store before the constraints).
  const std::vector<std::pair<std::string, Itv>>& new_vars() const { return fresh_; }

private:
  static bool pir_op(int s) {
    return s == ADD || s == MUL || s == TDIV || s == FDIV || s == CDIV || s == EDIV || s == MIN || s == MAX || s == EQ || s == LEQ;
  }
  static bool inf(int b) { return b == INT_MIN || b == INT_MAX; }
  static int clamp(long long v) { return v >= INT_MAX ? INT_MAX : v <= INT_MIN ? INT_MIN : (int)v; }

  std::string fresh(const char* prefix, Itv d) {
    std::string n = std::string("__") + prefix + std::to_string(counter_++);
    env_.declare(n);
    doms_[n] = d;
    fresh_.push_back({n, d});
    return n;
  }
  std::string konst(int k) {
    auto it = consts_.find(k);
    if(it != consts_.end()) return it->second;
    return consts_[k] = fresh("c", Itv(k, k));
  }
  // a unary constraint on one variable: goes to the store (PIR::interpret_tell's `x <op> k`), and into the hull bookkeeping
  bool bound(const std::string& x, int cmp, int k) {
    out_.push_back(F::binary(F::var(x), cmp, F::z(k)));
    Itv& d = doms_[x];
    if(cmp == LEQ) d.u = std::min(d.u, k);
    else if(cmp == GEQ) d.l = std::max(d.l, k);
    else if(cmp == LT) d.u = std::min(d.u, k == INT_MIN ? k : k - 1);
    else if(cmp == GT) d.l = std::max(d.l, k == INT_MAX ? k : k + 1);
    else { d.l = std::max(d.l, k); d.u = std::min(d.u, k); }
    return true;
  }
  bool emit(int op, const std::string& x, const std::string& y, const std::string& z) {
    out_.push_back(F::binary(F::var(x), EQ, F::binary(F::var(y), op, F::var(z))));
    return true;
  }
  // initial domain of t = y op z (SUB = the difference y - z)
  Itv hull(int op, const std::string& ys, const std::string& zs) const {
    const Itv y = doms_.at(ys), z = doms_.at(zs);
    long long lo, hi;
    switch(op) {
      case ADD: lo = (inf(y.l) || inf(z.l)) ? INT_MIN : (long long)y.l + z.l; hi = (inf(y.u) || inf(z.u)) ? INT_MAX : (long long)y.u + z.u; break;
      case SUB: lo = (inf(y.l) || inf(z.u)) ? INT_MIN : (long long)y.l - z.u; hi = (inf(y.u) || inf(z.l)) ? INT_MAX : (long long)y.u - z.l; break;
      case MUL: {
        if(inf(y.l) || inf(y.u) || inf(z.l) || inf(z.u)) return Itv::top();
        const long long p[4] = {(long long)y.l * z.l, (long long)y.l * z.u, (long long)y.u * z.l, (long long)y.u * z.u};
        lo = std::min(std::min(p[0], p[1]), std::min(p[2], p[3])); hi = std::max(std::max(p[0], p[1]), std::max(p[2], p[3]));
        break;
      }
      case MIN: lo = std::min(y.l, z.l); hi = std::min(y.u, z.u); break;
      case MAX: lo = std::max(y.l, z.l); hi = std::max(y.u, z.u); break;
      default: {   // divisions: |y / z| <= |y|
        if(inf(y.l) || inf(y.u)) return Itv::top();
        const long long m = std::max(std::llabs((long long)y.l), std::llabs((long long)y.u));
        lo = -m; hi = m;
      }
    }
    return Itv(clamp(lo), clamp(hi));
  }
  // a variable t with t = y op z; identical definitions share one variable
  std::string define(int op, const std::string& y, const std::string& z, bool boolean) {
    const auto key = std::make_tuple(op, y, z);
    auto it = cse_.find(key);
    if(it != cse_.end()) return it->second;
    std::string t;
    if(op == SUB) { t = fresh("t", hull(SUB, y, z)); emit(ADD, y, t, z); }          // d = y - z  <=>  y = d + z
    else if(op == NOT) { t = fresh("b", Itv(0, 1)); emit(ADD, konst(1), t, y); }     // r = 1 - s  <=>  1 = r + s
    else { t = fresh(boolean ? "b" : "t", boolean ? Itv(0, 1) : hull(op, y, z)); emit(op, t, y, z); }
    return cse_[key] = t;
  }

  // a term -> the name of the variable that holds its value
  bool term(const TF& t, std::string& out) {
    if(t.is_variable()) { out = t.name; return doms_.count(t.name) != 0; }
    if(t.is_constant()) { out = konst(t.k); return true; }
    if(!t.is_seq()) return false;
    if(t.is_logical() || t.is_predicate()) return reify(t, out);   // a formula used as a 0/1 term (formula.hpp:1090-1102)
    const int sig = t.sig();
    std::string a, b;
    if(sig == NEG && t.args.size() == 1) { if(!term(t.seq(0), a)) return false; out = define(SUB, konst(0), a, false); return true; }
    if(sig == ABS && t.args.size() == 1) {
      if(!term(t.seq(0), a)) return false;
      const bool seen = cse_.count(std::make_tuple((int)MAX, a, define(SUB, konst(0), a, false))) != 0;
      out = define(MAX, a, define(SUB, konst(0), a, false), false);
      if(!seen) { emit(LEQ, konst(1), konst(0), out); Itv& d = doms_[out]; d.l = std::max(d.l, 0); }   // |a| >= 0
      return true;
    }
    if(t.args.size() < 2) return false;
    if(!term(t.seq(0), a)) return false;
    for(size_t i = 1; i < t.args.size(); ++i) {   // n-ary + and * as left-nested chains
      if(!term(t.seq((int)i), b)) return false;
      if(sig == SUB) a = define(SUB, a, b, false);
      else if(pir_op(sig) && sig != EQ && sig != LEQ) a = define(sig, a, b, false);
      else return false;
      if(t.args.size() > 2 && sig != ADD && sig != MUL) return false;
    }
    out = a;
    return true;
  }

  // a formula -> the name of a 0/1 variable equivalent to it
  bool reify(const TF& f, std::string& out) {
    if(f.is_variable()) { out = f.name; return doms_.count(f.name) != 0; }   // a Boolean variable is its own reification
    if(!f.is_seq()) return false;
    const int sig = f.sig();
    std::string a, b;
    if(sig == NOT && f.args.size() == 1) { if(!reify(f.seq(0), a)) return false; out = define(NOT, a, a, true); return true; }
    if((sig == AND || sig == OR) && f.args.size() >= 2) {
      if(!reify(f.seq(0), a)) return false;
      for(size_t i = 1; i < f.args.size(); ++i) { if(!reify(f.seq((int)i), b)) return false; a = define(sig == AND ? MIN : MAX, a, b, true); }
      out = a;
      return true;
    }
    if(!f.is_binary()) return false;
    const TF& l = f.seq(0); const TF& r = f.seq(1);
    const bool logical = l.is_logical() || l.is_predicate() || r.is_logical() || r.is_predicate();
    switch(sig) {
      case LEQ: if(!term(l, a) || !term(r, b)) return false; out = define(LEQ, a, b, true); return true;
      case GEQ: if(!term(l, a) || !term(r, b)) return false; out = define(LEQ, b, a, true); return true;
      case GT: if(!term(l, a) || !term(r, b)) return false; a = define(LEQ, a, b, true); out = define(NOT, a, a, true); return true;
      case LT: if(!term(l, a) || !term(r, b)) return false; a = define(LEQ, b, a, true); out = define(NOT, a, a, true); return true;
      case NEQ: if(!term(l, a) || !term(r, b)) return false; a = define(EQ, a, b, true); out = define(NOT, a, a, true); return true;
      case EQ:
        if(!logical) { if(!term(l, a) || !term(r, b)) return false; out = define(EQ, a, b, true); return true; }
        [[fallthrough]];   // `=` between formulas is `<=>`
      case EQUIV: if(!reify(l, a) || !reify(r, b)) return false; out = define(EQ, a, b, true); return true;
      case IMPLY: if(!reify(l, a) || !reify(r, b)) return false; out = define(LEQ, a, b, true); return true;
      case XOR: if(!reify(l, a) || !reify(r, b)) return false; a = define(EQ, a, b, true); out = define(NOT, a, a, true); return true;
      default: return false;
    }
  }

  VarEnv& env_;
  std::map<std::string, Itv> doms_;
  std::vector<F> out_;
  std::vector<std::pair<std::string, Itv>> fresh_;
  std::map<int, std::string> consts_;
  std::map<std::tuple<int, std::string, std::string>, std::string> cse_;
  int counter_ = 0;
};

} // namespace b200pc
