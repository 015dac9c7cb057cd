// b200pc/pir.hpp — header-only C++ façade over the C-ABI (include/lpc.h) that keeps the API surface of
// lala::PIR<A, Alloc> (lala-pc include/lala/pir.hpp) for the fixpoint hot path, so that code written against the
// reference (its tests, a solver's propagation loop) ports by changing the namespace.
//
// What is mirrored (reference line numbers in brackets):
//   bytecode_type {op, x, y, z}                     [pir.hpp:37-47]
//   PIR(AType, sub_ptr)                             [pir.hpp:158-163]   the store handle plays the role of sub_ptr
//   PIR(const PIR&, AbstractDeps&)                  [pir.hpp:182-205]   shared table, cloned store
//   interpret_tell / interpret_ask / deduce(tell)   [pir.hpp:254-352]   host side: formula -> bytecodes, sort, clamp
//   load_deduce(i), num_deductions()                [pir.hpp:358-384]
//   deduce(i), deduce(bytecode), ask(i)             [pir.hpp:368-390, 721-817]  one kernel launch each (parity use)
//   embed, is_bot, is_top, operator[], project, vars [pir.hpp:354-356, 831-855]
//   snapshot / restore                              [pir.hpp:857-870]
//   is_extractable / extract                        [pir.hpp:873-898]
// and the loop driver the reference borrows from lala-core, GaussSeidelIteration::fixpoint(n, f[, has_changed])
// (call sites tests/pir_test.cpp:60-62, 82-86). `PIR::fixpoint()` is the fast path: the whole loop as one persistent
// CUDA kernel (lpc_fixpoint).
//
//   deinterpret(env[, remove_entailed, n])          [pir.hpp:901-941]   table + store back to a conjunction; the
//                                                                       entailed propagators found by ONE lpc_ask_bits call
// Not mirrored: allocators (device memory is owned by the C-ABI handles), the IDiagnostics tree (interpretation
// errors are returned as a string). There is no CPU implementation behind this header: every call ends
// in liblpc.so, and construction throws when no CUDA device is present.
#pragma once
#include <algorithm>
#include <climits>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/lpc.h"

namespace b200pc {

using AType = int;
constexpr AType UNTYPED = -1;

// lala-core Sig values of the operators PIR accepts (include/lpc.h).
enum Sig : int { ADD = LPC_ADD, MUL = LPC_MUL, MIN = LPC_MIN, MAX = LPC_MAX, TDIV = LPC_TDIV, FDIV = LPC_FDIV,
                 CDIV = LPC_CDIV, EDIV = LPC_EDIV, EQ = LPC_EQ, LEQ = LPC_LEQ,
                 GEQ = 1001, NEQ = 1002, LT = 1003, GT = 1004,     // unary-store comparisons only
                 AND = 1014 };                                      // the n-ary conjunction deinterpret produces

inline void check(int rc) {
  if(rc != LPC_OK) throw std::runtime_error(std::string("lpc: ") + lpc_last_error());
}

// (abstract type, variable id) packed in one int like lala-core's AVar (vid = value >> 8, aty = value & 0xFF).
struct AVar {
  int value = -1;
  AVar() = default;
  AVar(AType aty, int vid) : value((vid << 8) | (aty & 0xFF)) {}
  int vid() const { return value >> 8; }
  AType aty() const { return value & 0xFF; }
  bool operator==(const AVar& o) const { return value == o.value; }
};

// Interval<ZLB>: [lb, ub] over int32; top = [INT_MIN, INT_MAX]; any lb > ub is bot, all bots compare equal.
struct Itv {
  struct Bound { int v; int value() const { return v; } };
  int l = INT_MIN, u = INT_MAX;
  Itv() = default;
  Itv(int lb, int ub) : l(lb), u(ub) {}
  explicit Itv(int k) : l(k), u(k) {}
  static Itv top() { return Itv(); }
  static Itv bot() { return Itv(INT_MAX, INT_MIN); }
  static Itv eq_zero() { return Itv(0, 0); }
  static Itv eq_one() { return Itv(1, 1); }
  Bound lb() const { return Bound{l}; }
  Bound ub() const { return Bound{u}; }
  bool is_bot() const { return l > u; }
  bool is_top() const { return l == INT_MIN && u == INT_MAX; }
  bool operator==(const Itv& o) const { return (is_bot() && o.is_bot()) || (l == o.l && u == o.u); }
  bool operator!=(const Itv& o) const { return !(*this == o); }
  // lattice order of lala-core >= v1.2: a >= b  <=>  a contains b
  bool operator>=(const Itv& o) const { return o.is_bot() || (l <= o.l && u >= o.u); }
};

// The constraint X = Y op Z (pir.hpp:37-47).
struct bytecode_type {
  Sig op;
  AVar x, y, z;
  const AVar& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

// VStore<Interval<ZLB>> on the device.
class VStore {
public:
  explicit VStore(int nvars) : n_(nvars) { check(lpc_store_create(nvars, &h_)); }
  VStore(const VStore& o) : n_(o.n_) { check(lpc_store_create(n_, &h_)); check(lpc_store_copy(h_, o.h_)); }
  VStore& operator=(const VStore&) = delete;
  ~VStore() { lpc_store_destroy(h_); }
  int vars() const { return n_; }
  Itv operator[](int v) const { int b[2]; check(lpc_store_read(h_, v, 1, b)); return Itv(b[0], b[1]); }
  Itv project(AVar x) const { return (*this)[x.vid()]; }
  bool embed(AVar x, const Itv& u) { int c = 0; check(lpc_store_embed(h_, x.vid(), u.l, u.u, &c)); return c != 0; }
  bool is_bot() const { int b = 0; check(lpc_store_is_bot(h_, &b)); return b != 0; }
  bool is_top() const { int b = 0; check(lpc_store_is_top(h_, &b)); return b != 0; }
  std::vector<int> snapshot() const { std::vector<int> s(2 * (size_t)n_); if(n_) check(lpc_store_read(h_, 0, n_, s.data())); return s; }
  void restore(const std::vector<int>& s) { if(n_) check(lpc_store_write(h_, 0, n_, s.data())); }
  void grow(int nvars) {   // new variables start at top
    if(nvars <= n_) return;
    lpc_store* g = nullptr;
    check(lpc_store_create(nvars, &g));
    std::vector<int> s = snapshot();
    if(n_) check(lpc_store_write(g, 0, n_, s.data()));
    lpc_store_destroy(h_);
    h_ = g; n_ = nvars;
  }
  lpc_store* handle() const { return h_; }
private:
  lpc_store* h_ = nullptr;
  int n_ = 0;
};

// Copy policy of the reference's AbstractDeps: shared copies reuse the root's propagator table (pir.hpp:182-195).
struct AbstractDeps {
  bool shared_copy = true;
  bool is_shared_copy() const { return shared_copy; }
};

// A tiny formula type: just enough structure for PIR::interpret_formula (pir.hpp:254-287) — variables, integer
// constants and binary nodes — standing in for lala-core's TFormula.
struct F {
  enum Kind { VAR, CONST, BINARY, NARY } kind = CONST;
  std::string name;
  int k = 0;
  int sig_ = 0;
  std::shared_ptr<F> a, b;
  std::vector<F> children;   // NARY (the conjunctions deinterpret builds, pir.hpp:912-925)
  static F nary(int sig, std::vector<F> seq) { F f; f.kind = NARY; f.sig_ = sig; f.children = std::move(seq); return f; }
  bool is_nary() const { return kind == NARY; }
  size_t arity() const { return kind == NARY ? children.size() : (kind == BINARY ? 2 : 0); }
  const F& child(size_t i) const { return kind == NARY ? children[i] : seq((int)i); }
  // a printable form, for diagnostics and tests: (x == (y + z)), (and ...)
  std::string str() const {
    auto op = [](int s) -> std::string {
      switch(s) {
        case LPC_ADD: return "+"; case LPC_MUL: return "*"; case LPC_MIN: return "min"; case LPC_MAX: return "max";
        case LPC_TDIV: return "tdiv"; case LPC_FDIV: return "fdiv"; case LPC_CDIV: return "cdiv"; case LPC_EDIV: return "ediv";
        case LPC_EQ: return "=="; case LPC_LEQ: return "<="; case 1001: return ">="; case 1002: return "!="; case 1003: return "<";
        case 1004: return ">"; case 1014: return "and"; default: return "op" + std::to_string(s);
      }
    };
    if(kind == VAR) return name;
    if(kind == CONST) return std::to_string(k);
    if(kind == BINARY) return "(" + a->str() + " " + op(sig_) + " " + b->str() + ")";
    std::string r = "(" + op(sig_);
    for(auto& c : children) r += " " + c.str();
    return r + ")";
  }
  static F var(const std::string& n) { F f; f.kind = VAR; f.name = n; return f; }
  static F z(int k) { F f; f.kind = CONST; f.k = k; return f; }
  static F binary(const F& l, int sig, const F& r) {
    F f; f.kind = BINARY; f.sig_ = sig; f.a = std::make_shared<F>(l); f.b = std::make_shared<F>(r); return f;
  }
  bool is_binary() const { return kind == BINARY; }
  bool is_variable() const { return kind == VAR; }
  bool is_constant() const { return kind == CONST; }
  int sig() const { return sig_; }
  const F& seq(int i) const { return i == 0 ? *a : *b; }
};

// Name -> AVar environment (lala-core VarEnv): variables are numbered in declaration order.
class VarEnv {
public:
  explicit VarEnv(AType store_type = 0) : aty_(store_type) {}
  AVar declare(const std::string& name) {
    auto it = vars_.find(name);
    if(it != vars_.end()) return it->second;
    AVar v(aty_, (int)vars_.size());
    vars_[name] = v;
    names_.push_back(name);
    return v;
  }
  bool interpret(const F& f, AVar& out) const {
    if(!f.is_variable()) return false;
    auto it = vars_.find(f.name);
    if(it == vars_.end()) return false;
    out = it->second;
    return true;
  }
  int num_vars() const { return (int)vars_.size(); }
  const std::string& name_of(AVar v) const { return names_[v.vid()]; }
private:
  AType aty_;
  std::map<std::string, AVar> vars_;
  std::vector<std::string> names_;
};

// The bounds of one variable as formulas: `x == k` for a singleton, else `x >= lb` / `x <= ub` for each finite end
// (the shape of lala-core's VStore / Interval deinterpret is un-vendored; this is its logical content).
inline void deinterpret_bounds(const std::string& name, const Itv& d, std::vector<F>& seq) {
  if(d.l == d.u) { seq.push_back(F::binary(F::var(name), EQ, F::z(d.l))); return; }
  if(d.l != INT_MIN) seq.push_back(F::binary(F::var(name), GEQ, F::z(d.l)));
  if(d.u != INT_MAX) seq.push_back(F::binary(F::var(name), LEQ, F::z(d.u)));
}
// sub->deinterpret(env) (pir.hpp:916): the store as a conjunction over the variables the environment names.
inline F deinterpret_store(const VStore& s, const VarEnv& env) {
  std::vector<F> seq;
  const std::vector<int> snap = s.snapshot();
  const int n = std::min(env.num_vars(), s.vars());
  for(int v = 0; v < n; ++v) deinterpret_bounds(env.name_of(AVar(0, v)), Itv(snap[2 * (size_t)v], snap[2 * (size_t)v + 1]), seq);
  return F::nary(AND, std::move(seq));
}

// GaussSeidelIteration of lala-core (fixpoint.hpp): sequential sweeps until no deduction changes anything.
struct GaussSeidelIteration {
  template <class Fn> bool iterate(size_t n, const Fn& f) const {
    bool changed = false;
    for(size_t i = 0; i < n; ++i) changed |= f(i);
    return changed;
  }
  template <class Fn> size_t fixpoint(size_t n, const Fn& f, bool& has_changed) const {
    size_t it = 0;
    bool changed = true;
    while(changed) { changed = iterate(n, f); has_changed |= changed; ++it; }
    return it;
  }
  template <class Fn, class Stop> size_t fixpoint(size_t n, const Fn& f, const Stop& must_stop, bool& has_changed) const {
    size_t it = 0;
    bool changed = true;
    while(changed && !must_stop()) { changed = iterate(n, f); has_changed |= changed; ++it; }
    return it;
  }
  template <class Fn> size_t fixpoint(size_t n, const Fn& f) const { bool c = false; return fixpoint(n, f, c); }
};

struct fixpoint_stats { bool has_changed = false, is_bot = false; int sweeps = 0; long long deductions = 0; float device_ms = 0; };

class PIR {
public:
  using sub_type = VStore;
  using universe_type = Itv;
  using local_universe_type = Itv;
  using sub_ptr = std::shared_ptr<VStore>;
  static constexpr const char* name = "PIR";
  static constexpr bool is_abstract_universe = false;
  static constexpr bool sequential = false;   // deduce(i) for distinct i may run concurrently (atomic joins)
  static constexpr bool preserve_bot = true;

  struct tell_type {
    std::vector<std::pair<AVar, Itv>> sub_value;   // store tells
    std::vector<bytecode_type> bytecodes;
  };
  using ask_type = tell_type;
  struct snapshot_type { int num_bytecodes; std::vector<int> sub_snap; };

  PIR(AType atype, sub_ptr sub) : atype_(atype), sub_(std::move(sub)), table_(std::make_shared<Table>()) {}
  // pir.hpp:198-205: the store is cloned; the table is shared when deps.is_shared_copy().
  PIR(const PIR& other, AbstractDeps& deps)
    : atype_(other.atype_), sub_(std::make_shared<VStore>(*other.sub_)),
      table_(deps.is_shared_copy() ? other.table_ : std::make_shared<Table>(*other.table_)) {
    if(!deps.is_shared_copy()) table_->h = nullptr;
  }
  AType aty() const { return atype_; }

  // pir.hpp:290-323. Accepts `X = Y op Z` / `Y op Z = X` over three variables, or a unary bound `x <op> k` that the
  // store absorbs (what sub->interpret does in the reference). On failure returns false and sets `why`.
  bool interpret_tell(const F& f, VarEnv& env, tell_type& tell, std::string* why = nullptr) const { return interpret(f, env, tell, why); }
  bool interpret_ask(const F& f, const VarEnv& env, ask_type& ask, std::string* why = nullptr) const { return interpret(f, const_cast<VarEnv&>(env), ask, why); }

  // pir.hpp:326-352
  bool deduce(const tell_type& t) {
    bool has_changed = false;
    for(auto& sv : t.sub_value) has_changed |= sub_->embed(sv.first, sv.second);
    if(!t.bytecodes.empty()) {
      auto& bc = table_->bytecodes;
      for(auto& b : t.bytecodes) {
        bc.push_back(b);
        if(b.op == EQ || b.op == LEQ) sub_->embed(b.x, Itv(0, 1));
      }
      std::stable_sort(bc.begin(), bc.end(), [](const bytecode_type& a, const bytecode_type& b) {
        return a.op == b.op ? (a.y.vid() == b.y.vid() ? (a.x.vid() == b.x.vid() ? a.z.vid() < b.z.vid() : a.x.vid() < b.x.vid()) : a.y.vid() < b.y.vid()) : a.op < b.op;
      });
      // the device table grows incrementally: the new records are appended, sorted and merged in by the library with the
      // same stable (op, y, x, z) order, and only the part of the device image that moved is uploaded
      table_->append(t.bytecodes, sub_->vars());
      has_changed = true;
    }
    return has_changed;
  }

  bool embed(AVar x, const Itv& dom) { return sub_->embed(x, dom); }
  bytecode_type load_deduce(int i) const { return table_->bytecodes.at(i); }
  bytecode_type load_deductions(int i) const { return load_deduce(i); }   // spelling of BASELINE.json's north_star
  int num_deductions() const { return (int)table_->bytecodes.size(); }

  // pir.hpp:387-390: one propagator step on the device.
  bool deduce(int i) { int c = 0; check(lpc_deduce_one(table(), sub_->handle(), i, &c)); return c != 0; }
  bool ask(int i) const { int e = 0; check(lpc_ask_one(table(), sub_->handle(), i, &e)); return e != 0; }
  bool ask(const ask_type& t) const {
    // every bytecode of the query must be entailed; evaluated through a scratch table over the same store
    if(!t.bytecodes.empty()) {
      Table q; q.bytecodes = t.bytecodes;
      int64_t n = 0;
      check(lpc_ask_all(q.get(sub_->vars()), sub_->handle(), &n));
      if(n != (long long)t.bytecodes.size()) return false;
    }
    for(auto& sv : t.sub_value) if(!(sv.second >= (*sub_)[sv.first.vid()])) return false;
    return true;
  }

  // The whole GaussSeidelIteration::fixpoint(num_deductions(), deduce) loop as one persistent kernel.
  fixpoint_stats fixpoint(int mode = LPC_MODE_AUTO, int max_sweeps = 0, bool stop_on_bot = true) {
    lpc_fixpoint_opts o; lpc_fixpoint_default_opts(&o);
    o.mode = mode; o.max_sweeps = max_sweeps; o.stop_on_bot = stop_on_bot;
    lpc_fixpoint_result r;
    check(lpc_fixpoint(table(), sub_->handle(), &o, &r));
    fixpoint_stats s;
    s.has_changed = r.has_changed; s.is_bot = r.is_bot; s.sweeps = r.sweeps; s.deductions = r.deductions; s.device_ms = r.device_ms;
    return s;
  }

  bool is_bot() const { return sub_->is_bot(); }
  bool is_top() const { return sub_->is_top() && table_->bytecodes.empty(); }
  Itv operator[](int x) const { return (*sub_)[x]; }
  Itv project(AVar x) const { return sub_->project(x); }
  int vars() const { return sub_->vars(); }

  snapshot_type snapshot() const { return snapshot_type{num_deductions(), sub_->snapshot()}; }
  void restore(const snapshot_type& snap) {
    auto& bc = table_->bytecodes;
    if((int)bc.size() > snap.num_bytecodes) {
      // pir.hpp:865-868 pops from the back of the sorted table; the façade keeps that literal behaviour
      bc.resize(snap.num_bytecodes);
      table_->truncate(snap.num_bytecodes);
    }
    sub_->restore(snap.sub_snap);
  }

  bool is_extractable() const {
    if(is_bot()) return false;
    int64_t n = 0;
    check(lpc_ask_all(table(), sub_->handle(), &n));
    return n == num_deductions();
  }
  void extract(PIR& ua) const { ua.sub_->restore(sub_->snapshot()); }
  void extract(VStore& ua) const { ua.restore(sub_->snapshot()); }

  // pir.hpp:901-941. One propagator back to `X = Y op Z`; the whole element as AND(store, propagators...), optionally
  // without the propagators PIR::ask entails (their count is added to num_entailed) - found with ONE lpc_ask_bits call
  // over the table instead of num_deductions() single asks.
  F deinterpret(const bytecode_type& b, const VarEnv& env) const {
    return F::binary(F::var(env.name_of(b.x)), EQ, F::binary(F::var(env.name_of(b.y)), (int)b.op, F::var(env.name_of(b.z))));
  }
  F deinterpret(const VarEnv& env, bool remove_entailed, size_t& num_entailed) const {
    std::vector<F> seq;
    seq.push_back(deinterpret_store(*sub_, env));
    const auto& bc = table_->bytecodes;
    std::vector<uint8_t> ent(bc.size(), 0);
    if(remove_entailed && !bc.empty()) check(lpc_ask_bits(table(), sub_->handle(), ent.data()));
    for(size_t i = 0; i < bc.size(); ++i) {
      if(remove_entailed && ent[i]) { ++num_entailed; continue; }
      seq.push_back(deinterpret(bc[i], env));
    }
    return F::nary(AND, std::move(seq));
  }
  F deinterpret(const VarEnv& env) const { size_t n = 0; return deinterpret(env, false, n); }
  F deinterpret(const tell_type& t, const VarEnv& env) const {   // an intermediate (tell / ask) element, pir.hpp:931-940
    std::vector<F> sub, seq;
    for(auto& sv : t.sub_value) deinterpret_bounds(env.name_of(sv.first), sv.second, sub);
    seq.push_back(F::nary(AND, std::move(sub)));
    for(auto& b : t.bytecodes) seq.push_back(deinterpret(b, env));
    return F::nary(AND, std::move(seq));
  }

  sub_ptr sub() const { return sub_; }
  // table bytes copied to the device so far (diagnostic: incremental tells upload only what changed)
  long long table_uploaded_bytes() const { return table_->uploaded_bytes(); }

private:
  struct Table {
    std::vector<bytecode_type> bytecodes;
    lpc_table* h = nullptr;
    int h_nvars = -1;
    Table() = default;
    Table(const Table& o) : bytecodes(o.bytecodes) {}
    ~Table() { lpc_table_destroy(h); }
    void invalidate() { lpc_table_destroy(h); h = nullptr; }
    static std::vector<lpc_bytecode> raw(const std::vector<bytecode_type>& v) {
      std::vector<lpc_bytecode> r(v.size());
      for(size_t i = 0; i < r.size(); ++i) r[i] = lpc_bytecode{(int)v[i].op, v[i].x.vid(), v[i].y.vid(), v[i].z.vid()};
      return r;
    }
    // PIR::deduce(tell), pir.hpp:326-352: lpc_table_append + lpc_table_finalize(sort) on the live handle
    void append(const std::vector<bytecode_type>& fresh, int nvars) {
      if(!h) return;   // not on the device yet: get() uploads everything
      if(nvars > h_nvars) { check(lpc_table_set_nvars(h, nvars)); h_nvars = nvars; }
      std::vector<lpc_bytecode> r = raw(fresh);
      check(lpc_table_append(h, r.data(), (int64_t)r.size()));
      check(lpc_table_finalize(h, 1));
    }
    void truncate(int n) {
      if(!h) return;
      check(lpc_table_truncate(h, n));
      check(lpc_table_finalize(h, 1));
    }
    lpc_table* get(int nvars) {
      if(!h) {
        std::vector<lpc_bytecode> r = raw(bytecodes);
        check(lpc_table_create_empty(nvars, &h));
        check(lpc_table_append(h, r.data(), (int64_t)r.size()));
        check(lpc_table_finalize(h, 1));
        h_nvars = nvars;
      }
      else if(nvars > h_nvars) {   // the store grew (a tell declared variables)
        check(lpc_table_set_nvars(h, nvars));
        check(lpc_table_finalize(h, 1));
        h_nvars = nvars;
      }
      return h;
    }
    long long uploaded_bytes() const { return h ? (long long)lpc_table_uploaded_bytes(h) : 0; }
  };
  lpc_table* table() const { return table_->get(sub_->vars()); }

  static bool is_pir_op(int s) {
    return s == ADD || s == MUL || s == TDIV || s == FDIV || s == CDIV || s == EDIV || s == MIN || s == MAX || s == EQ || s == LEQ;
  }
  bool interpret(const F& f, VarEnv& env, tell_type& out, std::string* why) const {
    auto fail = [&](const char* m) { if(why) *why = m; return false; };
    if(!f.is_binary()) return fail("The shape of this formula is not supported.");
    // unary bound absorbed by the store: x <op> k
    if(f.seq(0).is_variable() && f.seq(1).is_constant()) {
      AVar x;
      if(!env.interpret(f.seq(0), x)) return fail("Could not interpret the variables in the environment.");
      int k = f.seq(1).k;
      switch(f.sig()) {
        case LEQ: out.sub_value.push_back({x, Itv(INT_MIN, k)}); return true;
        case GEQ: out.sub_value.push_back({x, Itv(k, INT_MAX)}); return true;
        case LT: out.sub_value.push_back({x, Itv(INT_MIN, k - 1)}); return true;
        case GT: out.sub_value.push_back({x, Itv(k + 1, INT_MAX)}); return true;
        case EQ: out.sub_value.push_back({x, Itv(k, k)}); return true;
        default: return fail("Uninterpretable formula in both PIR and its sub-domain.");
      }
    }
    int left = f.seq(0).is_binary() ? 1 : 0, right = f.seq(1).is_binary() ? 1 : 0;
    if(f.sig() == EQ && left + right == 1) {
      const F& X = f.seq(left);
      const F& Y = f.seq(right).seq(0);
      const F& Z = f.seq(right).seq(1);
      int op = f.seq(right).sig();
      if(X.is_variable() && Y.is_variable() && Z.is_variable() && is_pir_op(op)) {
        bytecode_type b;
        b.op = (Sig)op;
        if(env.interpret(X, b.x) && env.interpret(Y, b.y) && env.interpret(Z, b.z)) { out.bytecodes.push_back(b); return true; }
        return fail("Could not interpret the variables in the environment.");
      }
    }
    return fail("The shape of this formula is not supported.");
  }

  AType atype_;
  sub_ptr sub_;
  std::shared_ptr<Table> table_;
};

} // namespace b200pc
