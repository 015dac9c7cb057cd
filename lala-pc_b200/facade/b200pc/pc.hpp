// b200pc/pc.hpp — header-only C++ façade over the C-ABI of include/lpc_pc.h that keeps the API surface of
// lala::PC<A, Alloc> (lala-pc include/lala/pc.hpp) for the fixpoint hot path.
//
// What is mirrored (reference line numbers in brackets):
//   PC(AType, sub_ptr), copy with AbstractDeps         [pc.hpp:138-175]  shared propagators, cloned store
//   interpret_tell / interpret_ask                     [pc.hpp:217-623]  host side; sub-domain first, then formula
//   deduce(const tell_type&)                           [pc.hpp:625-645]  append, sort by (formula kind, length)
//   embed, ask(ask_type), ask(i), num_deductions       [pc.hpp:647-667]
//   load_deduce(i), deduce(i)                          [pc.hpp:669-680]  one kernel launch (parity use)
//   is_bot, is_top, operator[], project, vars          [pc.hpp:685-709]
//   snapshot / restore                                 [pc.hpp:711-723]
//   is_extractable / extract                           [pc.hpp:726-752]
//   deinterpret(env[, remove_entailed, n])             [pc.hpp:754-788]  each propagator back to the formula it was
//                                                                        interpreted from; entailed ones found by ONE
//                                                                        lpc_pc_ask_all call
// `PC::fixpoint()` is the fast path: GaussSeidelIteration::fixpoint(num_deductions(), deduce) (tests/pc_test.cpp:91-94,
// tests/pc_bitset_test.cpp:52-55) as one persistent CUDA kernel (lpc_pc_fixpoint / lpc_pc_fixpoint_bits).
//
// The store type selects the universe: PC<VStore> runs over Interval<ZLB> cells, PC<BitVStore> over NBitset<64> cells.
// Formula shapes with a flat device kind (include/lpc_pc.h) are flattened; every other shape over the node types PC
// has a device rule for (and / or / equiv / imply / xor nests, comparisons between arbitrary terms, + - * / min max neg
// abs, n-ary sums and products) keeps its tree as an LPC_PC_TREE propagator and is walked on the device (csrc/pc_tree.cuh). What is
// left - `in` over non-bitset stores, store-typed sub-formulas, trees deeper than the device walks - fails interpretation
// with the reference's message ("The shape of this formula is not supported."). No CPU implementation lives behind
// this header.
#pragma once
#include <cstdint>

#include "pir.hpp"
#include "../../../include/lpc_pc.h"

namespace b200pc {

// Formula signatures PC understands beyond the PIR operators (numeric values are façade-local).
enum PcSig : int { SUB = 1010, NEG = 1011, ABS = 1012, NOT = 1013, /* AND = 1014 lives in pir.hpp's Sig */ OR = 1015, EQUIV = 1016, IN = 1017,
                   IMPLY = 1018, XOR = 1019 };

// NBitset<64, local_memory, unsigned long long> (pc_bitset_test.cpp:23): bit 0 = "<= -1", bit i = value i - 1,
// bit 63 = ">= 62".
struct NBit {
  struct Bound { int v; int value() const { return v; } };
  uint64_t bits = ~0ull;
  NBit() = default;
  NBit(int lb, int ub) : bits(lpc_nbit_range(lb, ub)) {}
  explicit NBit(int k) : bits(lpc_nbit_range(k, k)) {}
  static NBit raw(uint64_t b) { NBit r; r.bits = b; return r; }
  static NBit from_set(std::initializer_list<int> vs) { uint64_t b = 0; for(int v : vs) b |= lpc_nbit_range(v, v); return raw(b); }
  static NBit top() { return NBit(); }
  static NBit bot() { return raw(0); }
  static NBit eq_zero() { return NBit(0); }
  static NBit eq_one() { return NBit(1); }
  bool is_bot() const { return bits == 0; }
  bool is_top() const { return bits == ~0ull; }
  Bound lb() const { return Bound{(bits & 1) ? INT_MIN : (bits ? __builtin_ctzll(bits) - 1 : INT_MAX)}; }
  Bound ub() const { return Bound{(bits >> 63) ? INT_MAX : (bits ? 62 - __builtin_clzll(bits) : INT_MIN)}; }
  NBit complement() const { return raw(~bits); }
  bool operator==(const NBit& o) const { return bits == o.bits; }
  bool operator!=(const NBit& o) const { return bits != o.bits; }
  bool operator>=(const NBit& o) const { return (o.bits & ~bits) == 0; }
};

// VStore<NBitset<64>> on the device: the 8-byte cells of an lpc_store read as one uint64 per variable.
class BitVStore {
public:
  using universe_type = NBit;
  static constexpr bool bitset = true;
  explicit BitVStore(int nvars) : n_(nvars) {
    check(lpc_store_create(nvars, &h_));
    std::vector<uint64_t> top((size_t)nvars, ~0ull);
    if(nvars) check(lpc_store_write_bits(h_, 0, nvars, top.data()));
  }
  BitVStore(const BitVStore& o) : n_(o.n_) { check(lpc_store_create(n_, &h_)); check(lpc_store_copy(h_, o.h_)); }
  BitVStore& operator=(const BitVStore&) = delete;
  ~BitVStore() { lpc_store_destroy(h_); }
  int vars() const { return n_; }
  NBit operator[](int v) const { uint64_t b; check(lpc_store_read_bits(h_, v, 1, &b)); return NBit::raw(b); }
  NBit project(AVar x) const { return (*this)[x.vid()]; }
  bool embed(AVar x, const NBit& u) { int c = 0; check(lpc_store_embed_bits(h_, x.vid(), u.bits, &c)); return c != 0; }
  bool is_bot() const { int b = 0; check(lpc_store_is_bot_bits(h_, &b)); return b != 0; }
  bool is_top() const { int b = 0; check(lpc_store_is_top_bits(h_, &b)); return b != 0; }
  std::vector<uint64_t> snapshot() const { std::vector<uint64_t> s((size_t)n_); if(n_) check(lpc_store_read_bits(h_, 0, n_, s.data())); return s; }
  void restore(const std::vector<uint64_t>& s) { if(n_) check(lpc_store_write_bits(h_, 0, n_, s.data())); }
  lpc_store* handle() const { return h_; }
private:
  lpc_store* h_ = nullptr;
  int n_ = 0;
};

// An n-ary formula / term tree: a stand-in for lala-core's TFormula with just the structure PC::interpret reads
// (is_variable / is_constant / is(Seq) / sig / seq(i), pc.hpp:300-604).
struct TF {
  enum Kind { VAR, CONST, SEQ, SET, BOOL } kind = CONST;   // BOOL: the formulas true (k = 1) / false (k = 0)
  std::string name;
  int k = 0;
  int sig_ = 0;
  std::vector<TF> args;
  std::vector<int> set;   // the S of `x in S`
  static TF var(const std::string& n) { TF f; f.kind = VAR; f.name = n; return f; }
  static TF z(int k) { TF f; f.kind = CONST; f.k = k; return f; }
  static TF make_true() { TF f; f.kind = BOOL; f.k = 1; return f; }
  static TF make_false() { TF f; f.kind = BOOL; f.k = 0; return f; }
  bool is_true() const { return kind == BOOL && k != 0; }
  bool is_false() const { return kind == BOOL && k == 0; }
  static TF make_nary(int sig, std::vector<TF> a) { TF f; f.kind = SEQ; f.sig_ = sig; f.args = std::move(a); return f; }
  static TF make_unary(int sig, const TF& a) { return make_nary(sig, {a}); }
  static TF make_binary(const TF& l, int sig, const TF& r) { return make_nary(sig, {l, r}); }
  static TF in(const TF& x, std::vector<int> s) { TF v; v.kind = SET; v.set = std::move(s); return make_binary(x, IN, v); }
  bool is_variable() const { return kind == VAR; }
  bool is_constant() const { return kind == CONST; }
  bool is_seq() const { return kind == SEQ; }
  bool is_binary() const { return kind == SEQ && args.size() == 2; }
  int sig() const { return sig_; }
  const TF& seq(int i) const { return args[i]; }
  bool is_logical() const { return kind == SEQ && (sig_ == AND || sig_ == OR || sig_ == EQUIV || sig_ == NOT || sig_ == IMPLY || sig_ == XOR); }
  bool is_predicate() const { return kind == SEQ && (sig_ == EQ || sig_ == NEQ || sig_ == LEQ || sig_ == GEQ || sig_ == LT || sig_ == GT || sig_ == IN); }
};

template <class S> struct universe_of { using type = Itv; static constexpr bool bitset = false; };
template <> struct universe_of<BitVStore> { using type = NBit; static constexpr bool bitset = true; };

template <class S>
class PC {
public:
  using sub_type = S;
  using universe_type = typename universe_of<S>::type;
  using local_universe_type = universe_type;
  using sub_ptr = std::shared_ptr<S>;
  static constexpr bool bitset = universe_of<S>::bitset;
  static constexpr const char* name = "PC";
  static constexpr bool is_abstract_universe = false;
  static constexpr bool sequential = false;
  static constexpr bool preserve_bot = true;

  // One interpreted propagator in the flat device encoding, with the reference's sort key: the index of its
  // pc::Formula alternative (formula.hpp:886-920) and its length() (formula.hpp:1134-1136).
  struct prop_type {
    int kind = 0, rhs = 0, bvar = -1;
    std::vector<lpc_pc_term> terms;
    int ref_kind = 0, length = 0;
    std::shared_ptr<TF> source;   // the formula this propagator was interpreted from (deinterpret, pc.hpp:754-788)
  };
  struct tell_type {
    std::vector<std::pair<AVar, universe_type>> sub_value;
    std::vector<prop_type> props;
  };
  using ask_type = tell_type;
  using sub_snap_type = decltype(std::declval<const S&>().snapshot());
  struct snapshot_type { int num_props; sub_snap_type sub_snap; };

  PC(AType atype, sub_ptr sub) : atype_(atype), sub_(std::move(sub)), table_(std::make_shared<Table>()) {}
  PC(const PC& other, AbstractDeps& deps)
    : atype_(other.atype_), sub_(std::make_shared<S>(*other.sub_)),
      table_(deps.is_shared_copy() ? other.table_ : std::make_shared<Table>(*other.table_)) {}
  AType aty() const { return atype_; }

  bool interpret_tell(const TF& f, VarEnv& env, tell_type& tell, std::string* why = nullptr) const { return interpret(f, env, tell, why); }
  bool interpret_ask(const TF& f, const VarEnv& env, ask_type& ask, std::string* why = nullptr) const { return interpret(f, const_cast<VarEnv&>(env), ask, why); }

  // pc.hpp:625-645
  bool deduce(const tell_type& t) {
    bool has_changed = false;
    for(auto& sv : t.sub_value) has_changed |= sub_->embed(sv.first, sv.second);
    if(!t.props.empty()) {
      auto& ps = table_->props;
      ps.insert(ps.end(), t.props.begin(), t.props.end());
      std::stable_sort(ps.begin(), ps.end(), [](const prop_type& a, const prop_type& b) {
        return a.ref_kind < b.ref_kind || (a.ref_kind == b.ref_kind && a.length < b.length);
      });
      table_->invalidate();
      return true;
    }
    return has_changed;
  }

  bool embed(AVar x, const universe_type& dom) { return sub_->embed(x, dom); }
  int num_deductions() const { return (int)table_->props.size(); }
  const prop_type& load_deduce(int i) const { return table_->props.at(i); }
  const prop_type& load_deductions(int i) const { return load_deduce(i); }

  // pc.hpp:671-680 / 661-663: one propagator step / entailment test on the device.
  bool deduce(int i) {
    int c = 0;
    if(bitset) check(lpc_pc_deduce_one_bits(table(), sub_->handle(), i, &c));
    else check(lpc_pc_deduce_one(table(), sub_->handle(), i, &c));
    return c != 0;
  }
  bool ask(int i) const {
    std::vector<uint8_t> bits((size_t)num_deductions() + 1);
    int64_t n = 0;
    ask_all(table(), &n, bits.data());
    return bits.at(i) != 0;
  }
  // pc.hpp:651-659
  bool ask(const ask_type& t) const {
    if(!t.props.empty()) {
      Table q; q.props = t.props;
      int64_t n = 0;
      ask_all(q.get(sub_->vars()), &n, nullptr);
      if(n != (int64_t)t.props.size()) return false;
    }
    for(auto& sv : t.sub_value) if(!(sv.second >= (*sub_)[sv.first.vid()])) return false;
    return true;
  }

  // The whole GaussSeidelIteration::fixpoint(num_deductions(), deduce, has_changed) loop as one persistent kernel.
  fixpoint_stats fixpoint(int max_sweeps = 0, bool stop_on_bot = true) {
    lpc_fixpoint_opts o; lpc_fixpoint_default_opts(&o);
    o.max_sweeps = max_sweeps; o.stop_on_bot = stop_on_bot;
    lpc_fixpoint_result r;
    if(bitset) check(lpc_pc_fixpoint_bits(table(), sub_->handle(), &o, &r));
    else check(lpc_pc_fixpoint(table(), sub_->handle(), &o, &r));
    fixpoint_stats s;
    s.has_changed = r.has_changed; s.is_bot = r.is_bot; s.sweeps = r.sweeps; s.deductions = r.deductions; s.device_ms = r.device_ms;
    return s;
  }

  bool is_bot() const { return sub_->is_bot(); }
  bool is_top() const { return sub_->is_top() && table_->props.empty(); }
  universe_type operator[](int x) const { return (*sub_)[x]; }
  universe_type project(AVar x) const { return sub_->project(x); }
  int vars() const { return sub_->vars(); }

  snapshot_type snapshot() const { return snapshot_type{num_deductions(), sub_->snapshot()}; }
  void restore(const snapshot_type& snap) {   // pc.hpp:716-723
    auto& ps = table_->props;
    if((int)ps.size() > snap.num_props) { ps.resize(snap.num_props); table_->invalidate(); }
    sub_->restore(snap.sub_snap);
  }

  bool is_extractable() const {   // pc.hpp:726-738
    if(is_bot()) return false;
    int64_t n = 0;
    ask_all(table(), &n, nullptr);
    return n == num_deductions();
  }
  void extract(PC& ua) const { ua.sub_->restore(sub_->snapshot()); }
  void extract(S& ua) const { ua.restore(sub_->snapshot()); }
  sub_ptr sub() const { return sub_; }

  // pc.hpp:754-788: AND(store, propagators...) - the store alone when there is no propagator -, optionally without the
  // propagators that are entailed (one lpc_pc_ask_all call over the table, their count added to num_entailed).
  TF deinterpret(const VarEnv& env, bool remove_entailed, size_t& num_entailed) const {
    TF subf = deinterpret_sub(env);
    const auto& ps = table_->props;
    if(ps.empty()) return subf;
    std::vector<TF> seq;
    seq.push_back(std::move(subf));
    std::vector<uint8_t> ent(ps.size() + 1, 0);
    if(remove_entailed) { int64_t n = 0; ask_all(table(), &n, ent.data()); }
    for(size_t i = 0; i < ps.size(); ++i) {
      if(remove_entailed && ent[i]) { ++num_entailed; continue; }
      seq.push_back(ps[i].source ? *ps[i].source : TF::z(1));
    }
    return TF::make_nary(AND, std::move(seq));
  }
  TF deinterpret(const VarEnv& env) const { size_t n = 0; return deinterpret(env, false, n); }
  TF deinterpret(const tell_type& t, const VarEnv& env) const {
    std::vector<TF> sub, seq;
    for(auto& sv : t.sub_value) deinterpret_domain(env.name_of(sv.first), sv.second, sub);
    seq.push_back(TF::make_nary(AND, std::move(sub)));
    for(auto& p : t.props) seq.push_back(p.source ? *p.source : TF::z(1));
    return TF::make_nary(AND, std::move(seq));
  }

private:
  // one domain as formulas: a singleton as `x == k`, an interval by its finite ends, a bitset with holes as `x in S`
  static void deinterpret_domain(const std::string& name, const universe_type& d, std::vector<TF>& seq) {
    const int lb = d.lb().value(), ub = d.ub().value();
    if constexpr(bitset) {
      const uint64_t full = lpc_nbit_range(lb, ub);
      if(d.bits != full && !(d.bits & 1) && !(d.bits >> 63)) {   // holes, all values finite: the set itself
        std::vector<int> vs;
        for(int b = 1; b < 63; ++b) if(d.bits >> b & 1) vs.push_back(b - 1);
        seq.push_back(TF::in(TF::var(name), vs));
        return;
      }
    }
    if(lb == ub) { seq.push_back(TF::make_binary(TF::var(name), EQ, TF::z(lb))); return; }
    if(lb != INT_MIN) seq.push_back(TF::make_binary(TF::var(name), GEQ, TF::z(lb)));
    if(ub != INT_MAX) seq.push_back(TF::make_binary(TF::var(name), LEQ, TF::z(ub)));
  }
  TF deinterpret_sub(const VarEnv& env) const {
    std::vector<TF> seq;
    const int n = std::min(env.num_vars(), sub_->vars());
    for(int v = 0; v < n; ++v) deinterpret_domain(env.name_of(AVar(0, v)), (*sub_)[v], seq);
    return TF::make_nary(AND, std::move(seq));
  }
  struct Table {
    std::vector<prop_type> props;
    lpc_pc_table* h = nullptr;
    int h_nvars = -1;
    Table() = default;
    Table(const Table& o) : props(o.props) {}
    ~Table() { lpc_pc_table_destroy(h); }
    void invalidate() { lpc_pc_table_destroy(h); h = nullptr; }
    lpc_pc_table* get(int nvars) {
      if(!h || h_nvars != nvars) {
        invalidate();
        std::vector<lpc_pc_prop> ps(props.size());
        std::vector<lpc_pc_term> ts;
        for(size_t i = 0; i < props.size(); ++i) {
          ps[i] = lpc_pc_prop{props[i].kind, (int32_t)ts.size(), (int32_t)props[i].terms.size(), props[i].rhs, props[i].bvar};
          ts.insert(ts.end(), props[i].terms.begin(), props[i].terms.end());
        }
        check(lpc_pc_table_create(ps.data(), (int64_t)ps.size(), ts.data(), (int64_t)ts.size(), nvars, &h));
        h_nvars = nvars;
      }
      return h;
    }
  };
  lpc_pc_table* table() const { return table_->get(sub_->vars()); }
  void ask_all(lpc_pc_table* t, int64_t* n, uint8_t* bits) const {
    if(bitset) check(lpc_pc_ask_all_bits(t, sub_->handle(), n, bits));
    else check(lpc_pc_ask_all(t, sub_->handle(), n, bits));
  }

  // ---- interpretation (pc.hpp:217-623) --------------------------------------------------------------------------------
  static bool fail(std::string* why, const char* m) { if(why) *why = m; return false; }

  // What the sub-domain absorbs (VStore::interpret in the reference): a bound on ONE variable. An interval store takes
  // x <op> k for <=, >=, <, >, =; a bitset store also takes x != k and x in S (pc_bitset_test.cpp:68-99).
  bool interpret_sub(const TF& f, VarEnv& env, tell_type& out) const {
    if(!f.is_binary() || !f.seq(0).is_variable()) return false;
    AVar x;
    if(!env.interpret(F::var(f.seq(0).name), x)) return false;
    if(f.sig() == IN && f.seq(1).kind == TF::SET) {
      if constexpr(bitset) {
        uint64_t b = 0;
        for(int v : f.seq(1).set) b |= lpc_nbit_range(v, v);
        out.sub_value.push_back({x, NBit::raw(b)});
        return true;
      }
      else if(f.seq(1).set.size() == 1) { out.sub_value.push_back({x, universe_type(f.seq(1).set[0], f.seq(1).set[0])}); return true; }
      return false;
    }
    if(!f.seq(1).is_constant()) return false;
    const int k = f.seq(1).k;
    switch(f.sig()) {
      case LEQ: out.sub_value.push_back({x, universe_type(INT_MIN, k)}); return true;
      case GEQ: out.sub_value.push_back({x, universe_type(k, INT_MAX)}); return true;
      case LT: out.sub_value.push_back({x, universe_type(INT_MIN, k - 1)}); return true;
      case GT: out.sub_value.push_back({x, universe_type(k + 1, INT_MAX)}); return true;
      case EQ: out.sub_value.push_back({x, universe_type(k, k)}); return true;
      case NEQ:
        if constexpr(bitset) { out.sub_value.push_back({x, NBit(k).complement()}); return true; }
        return false;
      default: return false;
    }
  }

  // A term that is a variable, constant * variable, or a flat sum of those (pc.hpp:300-330, 262-296) -> {coef, var};
  // `len` = Term::length() of the tree the reference would build (terms.hpp:41, 81, 433, 519-524).
  static bool linear_leaf(const TF& t, const VarEnv& env, lpc_pc_term& out, int& len) {
    AVar v;
    if(t.is_variable()) { if(!env.interpret(F::var(t.name), v)) return false; out = lpc_pc_term{1, v.vid()}; len = 1; return true; }
    if(t.is_binary() && t.sig() == MUL && t.seq(0).is_constant() && t.seq(1).is_variable() && t.seq(0).k != 0) {
      if(!env.interpret(F::var(t.seq(1).name), v)) return false;
      out = lpc_pc_term{t.seq(0).k, v.vid()}; len = 3; return true;
    }
    if(t.is_seq() && t.sig() == NEG && t.args.size() == 1 && t.seq(0).is_variable()) {   // Unary<Neg>: the bounds of -1 * x
      if(!env.interpret(F::var(t.seq(0).name), v)) return false;
      out = lpc_pc_term{-1, v.vid()}; len = 2; return true;
    }
    return false;
  }
  static bool linear(const TF& t, const VarEnv& env, std::vector<lpc_pc_term>& out, int& len) {
    lpc_pc_term leaf; int l = 0;
    if(linear_leaf(t, env, leaf, l)) { out.push_back(leaf); len = l; return true; }
    if(t.is_seq() && t.sig() == ADD && t.args.size() >= 2) {
      len = 1;
      for(auto& a : t.args) { if(!linear_leaf(a, env, leaf, l)) return false; out.push_back(leaf); len += l; }
      return true;
    }
    if(t.is_binary() && t.sig() == SUB && t.seq(1).is_variable()) {   // Binary<GroupSub>(leaf, Variable) == leaf + (-1) * y
      AVar v;
      if(!linear_leaf(t.seq(0), env, leaf, l) || !env.interpret(F::var(t.seq(1).name), v)) return false;
      out.push_back(leaf); out.push_back(lpc_pc_term{-1, v.vid()});
      len = 1 + l + 1;
      return true;
    }
    return false;
  }
  // A comparison of a linear term with a constant. `>=` and `<` are the flipped `<=` and `>` (pc.hpp:563-566).
  static bool lin_cmp(const TF& f, const VarEnv& env, prop_type& p) {
    if(!f.is_binary()) return false;
    int sig = f.sig();
    const TF* l = &f.seq(0); const TF* r = &f.seq(1);
    if(sig == GEQ) { sig = LEQ; std::swap(l, r); }
    else if(sig == LT) { sig = GT; std::swap(l, r); }
    if(sig != LEQ && sig != GT) return false;
    const bool const_left = l->is_constant();
    const TF* term = const_left ? r : l;
    const TF* cst = const_left ? l : r;
    if(!cst->is_constant() || term->is_constant()) return false;
    int len = 0;
    if(!linear(*term, env, p.terms, len)) return false;
    p.length = 1 + len + 1;
    if(sig == LEQ) { p.kind = const_left ? LPC_PC_LIN_GE : LPC_PC_LIN_LE; p.rhs = cst->k; p.ref_kind = 4 /* ILeq */; }
    else if(!const_left) { p.kind = LPC_PC_LIN_GT; p.rhs = cst->k; p.ref_kind = 5 /* IGt */; }
    else {   // k > term: Inequality<neg>(Constant, term) embeds [-inf, k - 1] (formula.hpp:786-792)
      if(cst->k == INT_MIN) return false;
      p.kind = LPC_PC_LIN_LE; p.rhs = cst->k - 1; p.ref_kind = 5;
    }
    return true;
  }
  static bool lin_le(const TF& f, const VarEnv& env, prop_type& p) {   // the shape a reification accepts
    return lin_cmp(f, env, p) && p.kind == LPC_PC_LIN_LE && p.ref_kind == 4;
  }
  static bool literal(const TF& f, const VarEnv& env, lpc_pc_term& out) {
    AVar v;
    if(f.is_variable()) { if(!env.interpret(F::var(f.name), v)) return false; out = lpc_pc_term{1, v.vid()}; return true; }
    if(f.is_seq() && f.sig() == NOT && f.args.size() == 1 && f.seq(0).is_variable()) {
      if(!env.interpret(F::var(f.seq(0).name), v)) return false;
      out = lpc_pc_term{-1, v.vid()}; return true;
    }
    return false;
  }

  // ---- the general case: the tree itself as an LPC_PC_TREE stream (include/lpc_pc.h) -------------------------------
  // Token numbers of the stream; heights are checked against the device interpreter's limits (csrc/pc_tree.cuh).
  enum { TK_CONST = 1, TK_VAR = 2, TK_NEG = 3, TK_ABS = 4, TK_ADD = 5, TK_SUB = 6, TK_MUL = 7, TK_NARY_ADD = 8, TK_MIN = 9,
         TK_MAX = 10, TK_TDIV = 11, TK_FDIV = 12, TK_CDIV = 13, TK_EDIV = 14, TK_NARY_MUL = 15, FK_LIT = 20, FK_NLIT = 21, FK_LEQ = 22, FK_GT = 23, FK_EQ = 24, FK_NEQ = 25, FK_AND = 26, FK_OR = 27,
         FK_EQUIV = 28, FK_IMPLY = 29, FK_XOR = 30, FK_TRUE = 32, FK_FALSE = 33, TREE_TERM_DEPTH = 8, TREE_FORM_DEPTH = 6 };
  // interpret_term (pc.hpp:217-296): returns the height (0 = not a term), `len` = Term::length()
  static int tree_term(const TF& t, const VarEnv& env, std::vector<int>& w, int& len) {
    AVar v;
    if(t.is_variable()) { if(!env.interpret(F::var(t.name), v)) return 0; w.push_back(TK_VAR); w.push_back(v.vid()); len = 1; return 1; }
    if(t.is_constant()) { w.push_back(TK_CONST); w.push_back(t.k); len = 1; return 1; }
    if(!t.is_seq()) return 0;
    if((t.sig() == NEG || t.sig() == ABS) && t.args.size() == 1) {
      w.push_back(t.sig() == NEG ? TK_NEG : TK_ABS);
      int l = 0; const int h = tree_term(t.seq(0), env, w, l);
      len = 1 + l;
      return h ? 1 + h : 0;
    }
    int tok = 0;
    switch(t.sig()) { case ADD: tok = TK_ADD; break; case SUB: tok = TK_SUB; break; case MUL: tok = TK_MUL; break;
                      case MIN: tok = TK_MIN; break; case MAX: tok = TK_MAX; break; case TDIV: tok = TK_TDIV; break;
                      case FDIV: tok = TK_FDIV; break; case CDIV: tok = TK_CDIV; break; case EDIV: tok = TK_EDIV; break;
                      default: return 0; }
    if(t.args.size() == 2) {
      w.push_back(tok);
      int l0 = 0, l1 = 0;
      const int h0 = tree_term(t.seq(0), env, w, l0), h1 = h0 ? tree_term(t.seq(1), env, w, l1) : 0;
      len = 1 + l0 + l1;
      return (h0 && h1) ? 1 + std::max(h0, h1) : 0;
    }
    if((tok == TK_ADD || tok == TK_MUL) && t.args.size() > 2) {   // Nary<Add> / Nary<Mul> (pc.hpp:253-254)
      w.push_back(tok == TK_ADD ? TK_NARY_ADD : TK_NARY_MUL); w.push_back((int)t.args.size());
      int h = 0; len = 1;
      for(auto& a : t.args) { int l = 0; const int ha = tree_term(a, env, w, l); if(!ha) return 0; h = std::max(h, ha); len += l; }
      return 1 + h;
    }
    return 0;
  }
  // interpret_formula (pc.hpp:454-572): returns the height of the connective nest (0 = no device rule); `ref_kind` =
  // index of the pc::Formula alternative the reference builds (formula.hpp:886-899), `len` = its length().
  static int tree_formula(const TF& f, const VarEnv& env, std::vector<int>& w, int& ref_kind, int& len, bool negate = false) {
    AVar v;
    if(f.is_true() || f.is_false()) {   // pc.hpp:515-522: Formula::make_false / make_true (formula.hpp:169-239)
      const bool t = f.is_true() != negate;
      w.push_back(t ? FK_TRUE : FK_FALSE); ref_kind = t ? 2 : 3; len = 1; return 1;
    }
    if(f.is_variable()) {
      if(!env.interpret(F::var(f.name), v)) return 0;
      w.push_back(negate ? FK_NLIT : FK_LIT); w.push_back(v.vid()); ref_kind = negate ? 1 : 0; len = 1; return 1;
    }
    if(!f.is_seq()) return 0;
    if(f.sig() == NOT && f.args.size() == 1) {   // negation is pushed into a literal or a comparison (pc.hpp:487-520)
      const TF& g = f.seq(0);
      if(g.is_variable() || g.is_predicate() || g.kind == TF::BOOL) return tree_formula(g, env, w, ref_kind, len, !negate);
      return 0;
    }
    if(f.args.size() > 2 && (f.sig() == AND || f.sig() == OR) && !negate) {   // binarised to the right (pc.hpp:454-463)
      std::vector<TF> rest(f.args.begin() + 1, f.args.end());
      return tree_formula(TF::make_binary(f.seq(0), f.sig(), TF::make_nary(f.sig(), rest)), env, w, ref_kind, len);
    }
    if(!f.is_binary()) return 0;
    const TF& a = f.seq(0); const TF& b = f.seq(1);
    int sig = f.sig();
    if(sig == EQ && (a.is_predicate() || b.is_predicate() || a.is_logical() || b.is_logical())) sig = EQUIV;   // pc.hpp:553-557
    if(sig == AND || sig == OR || sig == EQUIV || sig == IMPLY || sig == XOR) {
      if(negate) return 0;
      w.push_back(sig == AND ? FK_AND : sig == OR ? FK_OR : sig == EQUIV ? FK_EQUIV : sig == IMPLY ? FK_IMPLY : FK_XOR);
      ref_kind = sig == AND ? 8 : sig == OR ? 9 : sig == EQUIV ? 10 : sig == IMPLY ? 11 : 12;
      int k0, k1, l0 = 0, l1 = 0;
      const int h0 = tree_formula(a, env, w, k0, l0), h1 = h0 ? tree_formula(b, env, w, k1, l1) : 0;
      len = 1 + l0 + l1;
      return (h0 && h1) ? 1 + std::max(h0, h1) : 0;
    }
    const TF* l = &a; const TF* r = &b;
    if(sig == GEQ) { sig = LEQ; std::swap(l, r); }        // pc.hpp:565-566
    else if(sig == LT) { sig = GT; std::swap(l, r); }
    if(negate) sig = sig == LEQ ? GT : sig == GT ? LEQ : sig == EQ ? NEQ : sig == NEQ ? EQ : 0;
    int tok = 0;
    switch(sig) { case LEQ: tok = FK_LEQ; ref_kind = 4; break; case GT: tok = FK_GT; ref_kind = 5; break;
                  case EQ: tok = FK_EQ; ref_kind = 6; break; case NEQ: tok = FK_NEQ; ref_kind = 7; break; default: return 0; }
    w.push_back(tok);
    int l0 = 0, l1 = 0;
    const int h0 = tree_term(*l, env, w, l0), h1 = h0 ? tree_term(*r, env, w, l1) : 0;
    len = 1 + l0 + l1;
    return (h0 && h1 && std::max(h0, h1) <= TREE_TERM_DEPTH) ? 1 : 0;
  }
  bool interpret_tree(const TF& f, VarEnv& env, tell_type& out, std::string* why) const {
    std::vector<int> w;
    prop_type p;
    const int h = tree_formula(f, env, w, p.ref_kind, p.length);
    if(h == 0 || h > TREE_FORM_DEPTH) return fail(why, "The shape of this formula is not supported.");
    if(w.size() % 2) w.push_back(0);
    p.kind = LPC_PC_TREE;
    for(size_t i = 0; i < w.size(); i += 2) p.terms.push_back(lpc_pc_term{w[i], w[i + 1]});
    out.props.push_back(p);
    return true;
  }

  bool interpret_formula(const TF& f, VarEnv& env, tell_type& out, std::string* why) const {
    const size_t n0 = out.props.size();
    bool ok = interpret_flat(f, env, out, why);
    if(!ok) {
      out.props.resize(n0);
      ok = interpret_tree(f, env, out, why);   // no flat kind: the tree itself goes to the device
    }
    if(ok) for(size_t i = n0; i < out.props.size(); ++i) out.props[i].source = std::make_shared<TF>(f);
    return ok;
  }

  bool interpret_flat(const TF& f, VarEnv& env, tell_type& out, std::string* why) const {
    prop_type p;
    lpc_pc_term lit;
    AVar x, y;
    // literals and (binarised, pc.hpp:454-463) disjunctions of literals
    if(literal(f, env, lit)) {
      p.kind = LPC_PC_CLAUSE; p.terms = {lit}; p.ref_kind = lit.coef > 0 ? 0 : 1; p.length = 1;
      out.props.push_back(p); return true;
    }
    if(f.is_seq() && f.sig() == OR && f.args.size() >= 2) {
      std::vector<const TF*> flat;
      const TF* cur = &f;
      for(;;) {   // a flat n-ary OR or a right-nested binary one: same propagator
        for(size_t i = 0; i + 1 < cur->args.size(); ++i) flat.push_back(&cur->args[i]);
        const TF& last = cur->args.back();
        if(last.is_seq() && last.sig() == OR && last.args.size() >= 2) cur = &last; else { flat.push_back(&last); break; }
      }
      for(const TF* a : flat) { if(!literal(*a, env, lit)) return fail(why, "The shape of this formula is not supported."); p.terms.push_back(lit); }
      p.kind = LPC_PC_CLAUSE; p.ref_kind = 9 /* IDisj */; p.length = 2 * (int)p.terms.size() - 1;
      out.props.push_back(p); return true;
    }
    if(!f.is_binary()) return fail(why, "The shape of this formula is not supported.");
    const TF& a = f.seq(0); const TF& b = f.seq(1);
    // b <=> (sum <= k): `EQUIV`, or `=` with a predicate operand (pc.hpp:551-555)
    if(f.sig() == EQUIV || (f.sig() == EQ && (a.is_predicate() || b.is_predicate() || a.is_logical() || b.is_logical()))) {
      const TF* lhs = &a; const TF* rhs = &b;
      if(!lhs->is_variable() && rhs->is_variable()) std::swap(lhs, rhs);
      if(lhs->is_variable() && !bitset && env.interpret(F::var(lhs->name), x) && lin_le(*rhs, env, p)) {
        p.kind = LPC_PC_REIF_LIN_LE; p.bvar = x.vid(); p.ref_kind = 10 /* IBicond */; p.length = 1 + 1 + p.length;
        out.props.push_back(p); return true;
      }
      return fail(why, "The shape of this formula is not supported.");
    }
    if(f.sig() == EQ || f.sig() == NEQ) {
      const bool ne = f.sig() == NEQ;
      if(a.is_variable() && b.is_variable() && env.interpret(F::var(a.name), x) && env.interpret(F::var(b.name), y)) {
        p.kind = ne ? LPC_PC_NEQ : LPC_PC_EQ; p.terms = {{1, x.vid()}, {1, y.vid()}}; p.ref_kind = ne ? 7 : 6; p.length = 3;
        out.props.push_back(p); return true;
      }
      if(ne && a.is_variable() && b.is_constant() && env.interpret(F::var(a.name), x)) {
        p.kind = LPC_PC_NEQ; p.terms = {{1, x.vid()}}; p.rhs = b.k; p.ref_kind = 7; p.length = 3;
        out.props.push_back(p); return true;
      }
      if(!ne && a.is_seq() && a.sig() == ABS && a.args.size() == 1 && a.seq(0).is_variable() && b.is_variable() &&
         env.interpret(F::var(a.seq(0).name), x) && env.interpret(F::var(b.name), y)) {
        p.kind = LPC_PC_ABS_EQ; p.terms = {{1, x.vid()}, {1, y.vid()}}; p.ref_kind = 6; p.length = 4;
        out.props.push_back(p); return true;
      }
      int len = 0;
      if(!ne && !bitset && a.is_seq() && linear(a, env, p.terms, len)) {   // sum = k, sum = z
        if(b.is_constant()) { p.kind = LPC_PC_LIN_EQ; p.rhs = b.k; p.ref_kind = 6; p.length = 1 + len + 1; out.props.push_back(p); return true; }
        if(b.is_variable() && env.interpret(F::var(b.name), y)) {
          p.kind = LPC_PC_LIN_EQ_VAR; p.bvar = y.vid(); p.ref_kind = 6; p.length = 1 + len + 1; out.props.push_back(p); return true;
        }
      }
      return fail(why, "The shape of this formula is not supported.");
    }
    if(lin_cmp(f, env, p)) {
      if(bitset) return fail(why, "kept as a tree over a bitset store");   // interpret_formula falls back to interpret_tree
      out.props.push_back(p); return true;
    }
    return fail(why, "The shape of this formula is not supported.");
  }

  // pc.hpp:582-609: the sub-domain first, then PC's own formulas.
  bool interpret(const TF& f, VarEnv& env, tell_type& out, std::string* why) const {
    if(interpret_sub(f, env, out)) return true;
    return interpret_formula(f, env, out, why);
  }

  AType atype_;
  sub_ptr sub_;
  std::shared_ptr<Table> table_;
};

using IPC = PC<VStore>;        // tests/pc_test.cpp:26-28
using BitPC = PC<BitVStore>;   // tests/pc_bitset_test.cpp:23-25

} // namespace b200pc
