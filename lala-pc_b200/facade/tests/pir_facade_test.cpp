// pir_facade_test.cpp — the reference's PIR tests (lala-pc tests/pir_test.cpp, tests/bound_consistency_test.hpp)
// re-expressed against the b200pc façade. FlatZinc parsing and ternarisation are not part of the hot path, so models
// are built with F::binary(...) instead of FlatZinc strings; everything from interpret_tell on is the same call
// sequence the reference's helpers make: interpret_tell -> deduce(tell) -> GaussSeidelIteration::fixpoint(deduce(i))
// -> compare intervals -> is_extractable / extract. Needs a CUDA device (run by tests/test_gpu_facade.py).
#include <cstdio>
#include <cstdlib>
#include <initializer_list>
#include <string>
#include <vector>

#include "../b200pc/pir.hpp"

using namespace b200pc;

static int g_fail = 0, g_checks = 0;
#define EXPECT_TRUE(c) do { ++g_checks; if(!(c)) { ++g_fail; printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); } } while(0)
#define EXPECT_FALSE(c) EXPECT_TRUE(!(c))
#define EXPECT_EQ(a, b) do { ++g_checks; if(!((a) == (b))) { ++g_fail; printf("FAIL %s:%d: %s == %s\n", __FILE__, __LINE__, #a, #b); } } while(0)

using IStore = VStore;
using IPIR = PIR;
const AType sty = 0;
const AType pty = 1;

// A model under construction: declared variables, their domains and the constraints told so far.
struct Model {
  VarEnv env{sty};
  std::vector<std::pair<std::string, Itv>> doms;
  std::vector<F> cons;
  Model& var(const char* n, Itv d = Itv::top()) { env.declare(n); doms.push_back({n, d}); return *this; }
  Model& c(const F& f) { cons.push_back(f); return *this; }
};

// create_and_interpret_and_tell<IPIR, true>(...) of lala-core's abstract_testing.hpp
static IPIR create_and_interpret_and_tell(Model& m) {
  IPIR pir(pty, std::make_shared<IStore>(m.env.num_vars()));
  IPIR::tell_type tell;
  for(auto& d : m.doms) { AVar v; m.env.interpret(F::var(d.first), v); tell.sub_value.push_back({v, d.second}); }
  std::string why;
  for(auto& f : m.cons) EXPECT_TRUE(pir.interpret_tell(f, m.env, tell, &why));
  pir.deduce(tell);
  return pir;
}

// interpret_must_succeed<IKind::TELL>(...)
static void tell_more(IPIR& pir, Model& m, const F& f) {
  IPIR::tell_type tell;
  std::string why;
  EXPECT_TRUE(pir.interpret_tell(f, m.env, tell, &why));
  pir.deduce(tell);
}

// tests/pir_test.cpp:31-52
static void test_extract(const IPIR& pir, bool is_ua) {
  AbstractDeps deps;
  IPIR copy1(pir, deps);
  if(is_ua) for(int i = 0; i < pir.num_deductions(); ++i) EXPECT_TRUE(pir.ask(i));
  EXPECT_EQ(pir.is_extractable(), is_ua);
  if(pir.is_extractable()) {
    pir.extract(copy1);
    EXPECT_EQ(pir.is_top(), copy1.is_top());
    EXPECT_EQ(pir.is_bot(), copy1.is_bot());
    for(int i = 0; i < pir.vars(); ++i) EXPECT_EQ(pir[i], copy1[i]);
  }
}

// tests/pir_test.cpp:54-68, once with the reference's own loop (deduce(i) one launch at a time) and once with the
// fused device fixpoint, which must agree.
static void deduce_and_test(IPIR& pir, int num_deds, const std::vector<Itv>& before, const std::vector<Itv>& after, bool is_ua) {
  EXPECT_EQ(pir.num_deductions(), num_deds);
  for(size_t i = 0; i < before.size(); ++i) EXPECT_EQ(pir[i], before[i]);
  AbstractDeps deps;
  IPIR fused(pir, deps);
  GaussSeidelIteration{}.fixpoint(pir.num_deductions(), [&](size_t i) { return pir.deduce((int)i); });
  for(size_t i = 0; i < after.size(); ++i) EXPECT_EQ(pir[i], after[i]);
  test_extract(pir, is_ua);
  fixpoint_stats st = fused.fixpoint();
  EXPECT_FALSE(st.is_bot);
  for(int i = 0; i < pir.vars(); ++i) EXPECT_EQ(fused[i], pir[i]);
}
static void deduce_and_test(IPIR& pir, int num_deds, const std::vector<Itv>& before_after, bool is_ua = false) {
  deduce_and_test(pir, num_deds, before_after, before_after, is_ua);
}

// tests/pir_test.cpp:75-89
static void deduce_and_test_bot(IPIR& pir, int num_deds, const std::vector<Itv>& before) {
  EXPECT_EQ(pir.num_deductions(), num_deds);
  for(size_t i = 0; i < before.size(); ++i) EXPECT_EQ(pir[i], before[i]);
  AbstractDeps deps;
  IPIR fused(pir, deps);
  bool has_changed = false;
  // the reference iterates without a stop condition; its failed stores keep moving, so stop at bot (must_stop overload)
  GaussSeidelIteration{}.fixpoint(pir.num_deductions(), [&](size_t i) { return pir.deduce((int)i); },
                                  [&]() { return pir.is_bot(); }, has_changed);
  EXPECT_TRUE(has_changed);
  EXPECT_TRUE(pir.is_bot());
  fixpoint_stats st = fused.fixpoint();
  EXPECT_TRUE(st.is_bot && st.has_changed && fused.is_bot());
}

static F V(const char* n) { return F::var(n); }
static F X_eq(const char* x, const char* y, int op, const char* z) { return F::binary(V(x), EQ, F::binary(V(y), op, V(z))); }

static void TernaryProblem() {   // pir_test.cpp:175-182
  Model m;
  m.var("x", Itv(0, 10)).var("y", Itv(0, 10)).var("z", Itv(5, 5)).c(X_eq("z", "x", ADD, "y"));
  IPIR pir = create_and_interpret_and_tell(m);
  deduce_and_test(pir, 1, {Itv(0, 10), Itv(0, 10), Itv(5, 5)}, {Itv(0, 5), Itv(0, 5), Itv(5, 5)}, false);
}

static void TemporalConstraint1Flat() {   // pir_test.cpp:217-224: x + y = z, z <= 5
  Model m;
  m.var("x", Itv(0, 10)).var("y", Itv(0, 10)).var("z").c(F::binary(V("z"), LEQ, F::z(5))).c(X_eq("z", "x", ADD, "y"));
  IPIR pir = create_and_interpret_and_tell(m);
  deduce_and_test(pir, 1, {Itv(0, 10), Itv(0, 10), Itv(INT_MIN, 5)}, {Itv(0, 5), Itv(0, 5), Itv(0, 5)}, false);
}

static void TernaryAdds() {   // pir_test.cpp:322-359: t = x + y ; s = t + z ; s <= k
  struct { Itv d; int k; bool bot; Itv after; bool ua; } cases[] = {
    {Itv(3, 10), 8, true, Itv(), false}, {Itv(3, 10), 9, false, Itv(3, 3), true},
    {Itv(3, 10), 10, false, Itv(3, 4), false}, {Itv(-2, 2), -5, false, Itv(-2, -1), false}};
  for(auto& c : cases) {
    Model m;
    m.var("x", c.d).var("y", c.d).var("z", c.d).var("t").var("s");
    m.c(F::binary(V("s"), LEQ, F::z(c.k))).c(X_eq("t", "x", ADD, "y")).c(X_eq("s", "t", ADD, "z"));
    IPIR pir = create_and_interpret_and_tell(m);
    if(c.bot) deduce_and_test_bot(pir, 2, {c.d, c.d, c.d});
    else deduce_and_test(pir, 2, {c.d, c.d, c.d}, {c.after, c.after, c.after}, c.ua);
  }
}

static void MinConstraint1() {   // pir_test.cpp:626-640, incremental tells between fixpoints
  Model m;
  m.var("x", Itv(0, 4)).var("y", Itv(2, 5)).var("z", Itv(0, 10)).c(X_eq("z", "x", MIN, "y"));
  IPIR pir = create_and_interpret_and_tell(m);
  deduce_and_test(pir, 1, {Itv(0, 4), Itv(2, 5), Itv(0, 10)}, {Itv(0, 4), Itv(2, 5), Itv(0, 4)}, false);
  tell_more(pir, m, F::binary(V("z"), LEQ, F::z(3)));
  deduce_and_test(pir, 1, {Itv(0, 4), Itv(2, 5), Itv(0, 3)}, false);
  tell_more(pir, m, F::binary(V("x"), LEQ, F::z(1)));
  deduce_and_test(pir, 1, {Itv(0, 1), Itv(2, 5), Itv(0, 3)}, {Itv(0, 1), Itv(2, 5), Itv(0, 1)}, false);
  tell_more(pir, m, F::binary(V("x"), LEQ, F::z(0)));
  deduce_and_test(pir, 1, {Itv(0, 0), Itv(2, 5), Itv(0, 1)}, {Itv(0, 0), Itv(2, 5), Itv(0, 0)}, true);
}

static void MaxConstraint2() {   // pir_test.cpp:688-696
  Model m;
  m.var("x", Itv(0, 4)).var("y", Itv(2, 5)).var("z", Itv(0, 10)).c(X_eq("z", "x", MAX, "y"));
  IPIR pir = create_and_interpret_and_tell(m);
  deduce_and_test(pir, 1, {Itv(0, 4), Itv(2, 5), Itv(0, 10)}, {Itv(0, 4), Itv(2, 5), Itv(2, 5)}, false);
  tell_more(pir, m, F::binary(V("z"), GEQ, F::z(5)));
  deduce_and_test(pir, 1, {Itv(0, 4), Itv(2, 5), Itv(5, 5)}, {Itv(0, 4), Itv(5, 5), Itv(5, 5)}, true);
}

static void IntTimes1() {   // pir_test.cpp:759-771
  Model m;
  m.var("x", Itv(0, 1)).var("y", Itv(0, 1)).var("z", Itv(0, 1)).c(X_eq("z", "x", MUL, "y"));
  IPIR pir = create_and_interpret_and_tell(m);
  deduce_and_test(pir, 1, {Itv(0, 1), Itv(0, 1), Itv(0, 1)}, false);
  tell_more(pir, m, F::binary(V("x"), EQ, F::z(1)));
  deduce_and_test(pir, 1, {Itv(1, 1), Itv(0, 1), Itv(0, 1)}, false);
  tell_more(pir, m, F::binary(V("y"), EQ, F::z(1)));
  deduce_and_test(pir, 1, {Itv(1, 1), Itv(1, 1), Itv(0, 1)}, {Itv(1, 1), Itv(1, 1), Itv(1, 1)}, true);
}

static void IntDivs() {   // pir_test.cpp:875-913: x = y div z, x in 2..10, y in -25..25, z in -2..3
  struct { int op; int ylb; } cases[] = {{EDIV, -20}, {CDIV, -20}, {TDIV, -21}, {FDIV, -21}};
  for(auto& c : cases) {
    Model m;
    m.var("x", Itv(2, 10)).var("y", Itv(-25, 25)).var("z", Itv(-2, 3)).c(X_eq("x", "y", c.op, "z"));
    IPIR pir = create_and_interpret_and_tell(m);
    deduce_and_test(pir, 1, {Itv(2, 10), Itv(-25, 25), Itv(-2, 3)}, {Itv(2, 10), Itv(c.ylb, 25), Itv(-2, 3)}, false);
  }
}

static void InfiniteDomains() {   // pir_test.cpp:924-946: b <=> (x <= 5) with a constant variable for 5
  for(int bval = 0; bval <= 1; ++bval) {
    Model m;
    m.var("x").var("b", Itv(0, 1)).var("five", Itv(5, 5)).c(X_eq("b", "x", LEQ, "five"));
    IPIR pir = create_and_interpret_and_tell(m);
    deduce_and_test(pir, 1, {Itv::top(), Itv(0, 1)}, false);
    tell_more(pir, m, F::binary(V("b"), EQ, F::z(bval)));
    if(bval) deduce_and_test(pir, 1, {Itv::top(), Itv(1, 1)}, {Itv(INT_MIN, 5), Itv(1, 1)}, true);
    else deduce_and_test(pir, 1, {Itv::top(), Itv(0, 0)}, {Itv(6, INT_MAX), Itv(0, 0)}, true);
  }
}

static void Strict1() {   // pir_test.cpp:502-506: x > y  ==  ZERO = (x <= y)
  Model m;
  m.var("x", Itv(1, 10)).var("y", Itv(10, 10)).var("zero", Itv(0, 0)).c(X_eq("zero", "x", LEQ, "y"));
  IPIR pir = create_and_interpret_and_tell(m);
  deduce_and_test_bot(pir, 1, {Itv(1, 10)});
}

static void InterpretationErrors() {   // pir.hpp:254-287: only X = Y op Z over three variables is a PIR constraint
  Model m;
  m.var("x").var("y").var("z");
  IPIR pir = create_and_interpret_and_tell(m);
  IPIR::tell_type tell;
  std::string why;
  EXPECT_FALSE(pir.interpret_tell(F::binary(V("x"), EQ, F::binary(V("y"), ADD, F::z(3))), m.env, tell, &why));
  EXPECT_EQ(why, std::string("The shape of this formula is not supported."));
  EXPECT_FALSE(pir.interpret_tell(F::binary(V("x"), EQ, F::binary(V("y"), ADD, V("nope"))), m.env, tell, &why));
  EXPECT_EQ(why, std::string("Could not interpret the variables in the environment."));
  EXPECT_FALSE(pir.interpret_tell(F::binary(V("x"), LEQ, F::binary(V("y"), ADD, V("z"))), m.env, tell, &why));
  EXPECT_TRUE(tell.bytecodes.empty());
  EXPECT_TRUE(pir.is_top());
}

// bound_consistency_test.hpp:155-225 on a reduced range, with snapshot / restore / embed like the reference:
// every interval triple in [-3,3]^3 for x = y + z must give the exact hull of the concrete solutions (or bot).
static void ExhaustiveAddWithSnapshots() {
  const int lo = -3, hi = 3;
  Model m;
  m.var("x", Itv(lo, hi)).var("y", Itv(lo, hi)).var("z", Itv(lo, hi)).c(X_eq("x", "y", ADD, "z"));
  IPIR a = create_and_interpret_and_tell(m);
  auto snap = a.snapshot();
  AVar xv(sty, 0), yv(sty, 1), zv(sty, 2);
  int cases = 0;
  for(int xl = lo; xl <= hi; ++xl) for(int xu = xl; xu <= hi; ++xu)
  for(int yl = lo; yl <= hi; ++yl) for(int yu = yl; yu <= hi; ++yu)
  for(int zl = lo; zl <= hi; zl += 2) for(int zu = zl; zu <= hi; zu += 3) {
    a.restore(snap);
    a.embed(xv, Itv(xl, xu)); a.embed(yv, Itv(yl, yu)); a.embed(zv, Itv(zl, zu));
    Itv x2 = Itv::bot(), y2 = Itv::bot(), z2 = Itv::bot();
    auto join = [](Itv& h, int v) { if(h.is_bot()) h = Itv(v, v); else { h.l = std::min(h.l, v); h.u = std::max(h.u, v); } };
    for(int p = xl; p <= xu; ++p) for(int q = yl; q <= yu; ++q) for(int r = zl; r <= zu; ++r)
      if(p == q + r) { join(x2, p); join(y2, q); join(z2, r); }
    fixpoint_stats st = a.fixpoint();
    if(st.is_bot) EXPECT_TRUE(x2.is_bot());
    else { EXPECT_EQ(a[0], x2); EXPECT_EQ(a[1], y2); EXPECT_EQ(a[2], z2); }
    ++cases;
  }
  EXPECT_TRUE(cases > 1000);
}

// pir.hpp:326-352 + 857-870: CONSTRAINTS told between fixpoints (the table grows on the device without being rebuilt),
// snapshot / restore across a table change (restore pops the records told since), shared and private table copies.
static void IncrementalTells() {
  Model m;
  for(int i = 0; i < 40; ++i) m.var(("v" + std::to_string(i)).c_str(), Itv(0, 1000000000));
  m.c(X_eq("v2", "v0", ADD, "v1"));
  IPIR pir = create_and_interpret_and_tell(m);
  pir.embed(AVar(sty, 0), Itv(3, 5)); pir.embed(AVar(sty, 1), Itv(10, 20));
  EXPECT_FALSE(pir.fixpoint().is_bot);
  EXPECT_EQ(pir[2], Itv(13, 25));
  const long long b0 = pir.table_uploaded_bytes();
  EXPECT_TRUE(b0 > 0);
  auto snap = pir.snapshot();
  // a chain v(i+2) = v(i) + v(i+1), one constraint per tell; every tell sorts into the live table
  for(int i = 1; i < 30; ++i)
    tell_more(pir, m, X_eq(("v" + std::to_string(i + 2)).c_str(), ("v" + std::to_string(i)).c_str(), ADD, ("v" + std::to_string(i + 1)).c_str()));
  EXPECT_EQ(pir.num_deductions(), 30);
  // sorted by (op, y, x, z): y = v0, v1, ... in order
  for(int i = 0; i < 30; ++i) { EXPECT_EQ(pir.load_deduce(i).y.vid(), i); EXPECT_EQ(pir.load_deduce(i).x.vid(), i + 2); }
  // 29 tells uploaded 29 small pieces, not 29 whole tables: far less than 29 x the 16-record granule of the first upload
  EXPECT_TRUE(pir.table_uploaded_bytes() - b0 <= 29 * 2 * 16 * 13);
  fixpoint_stats st = pir.fixpoint();
  EXPECT_FALSE(st.is_bot);
  EXPECT_EQ(pir[3], Itv(23, 45));     // v3 = v1 + v2
  EXPECT_EQ(pir[4], Itv(36, 70));     // v4 = v2 + v3
  EXPECT_EQ(pir[5], Itv(59, 115));    // v5 = v3 + v4
  // a mixed tell: other operators land in their own runs
  tell_more(pir, m, X_eq("v35", "v0", MAX, "v1"));
  tell_more(pir, m, X_eq("v36", "v0", MUL, "v1"));
  EXPECT_EQ(pir.num_deductions(), 32);
  EXPECT_EQ((int)pir.load_deduce(30).op, (int)MUL);   // ADD = 2 < MUL = 4 < MAX = 7
  EXPECT_EQ((int)pir.load_deduce(31).op, (int)MAX);
  // a private copy owns its table, a shared copy uses the same one
  AbstractDeps shared, priv; priv.shared_copy = false;
  IPIR c1(pir, shared), c2(pir, priv);
  EXPECT_EQ(c1.num_deductions(), 32); EXPECT_EQ(c2.num_deductions(), 32);
  // restore pops everything told after the snapshot (pir.hpp:863-870) and the store with it
  pir.restore(snap);
  EXPECT_EQ(pir.num_deductions(), 1);
  EXPECT_EQ(pir[2], Itv(13, 25));
  EXPECT_EQ(pir[3], Itv(0, 1000000000));
  EXPECT_FALSE(pir.fixpoint().has_changed);
  EXPECT_EQ(c2.num_deductions(), 32);   // the private copy is unaffected
  c2.fixpoint();
  EXPECT_EQ(c2[35], Itv(10, 20));       // max(v0, v1)
}

// pir.hpp:901-941
static void Deinterpret() {
  Model m;
  m.var("x", Itv(1, 1)).var("y", Itv(0, 10)).var("z", Itv(0, 10)).var("w").c(X_eq("z", "x", ADD, "y")).c(X_eq("w", "x", MUL, "x"));
  IPIR pir = create_and_interpret_and_tell(m);
  EXPECT_EQ(pir.deinterpret(m.env).str(),
            std::string("(and (and (x == 1) (y >= 0) (y <= 10) (z >= 0) (z <= 10)) (z == (x + y)) (w == (x * x)))"));
  pir.fixpoint();
  size_t ent = 0;
  F f = pir.deinterpret(m.env, true, ent);
  // w = x * x with x = 1 is entailed once w is 1; z = x + y is not (y, z still ranges)
  EXPECT_EQ(ent, (size_t)1);
  EXPECT_EQ(f.str(), std::string("(and (and (x == 1) (y >= 0) (y <= 9) (z >= 1) (z <= 10) (w == 1)) (z == (x + y)))"));
  IPIR::tell_type tell;
  std::string why;
  EXPECT_TRUE(pir.interpret_tell(F::binary(V("y"), LEQ, F::z(4)), m.env, tell, &why));
  EXPECT_TRUE(pir.interpret_tell(X_eq("w", "y", MIN, "z"), m.env, tell, &why));
  EXPECT_EQ(pir.deinterpret(tell, m.env).str(), std::string("(and (and (y <= 4)) (w == (y min z)))"));
  // what deinterpret gives back can be told again: same element
  Model m2;
  m2.var("x").var("y").var("z").var("w");
  IPIR again(pty, std::make_shared<IStore>(4));
  IPIR::tell_type t2;
  F all = pir.deinterpret(m.env);
  for(size_t i = 0; i < all.child(0).arity(); ++i) EXPECT_TRUE(again.interpret_tell(all.child(0).child(i), m2.env, t2, &why));
  for(size_t i = 1; i < all.arity(); ++i) EXPECT_TRUE(again.interpret_tell(all.child(i), m2.env, t2, &why));
  again.deduce(t2);
  EXPECT_EQ(again.num_deductions(), pir.num_deductions());
  for(int v = 0; v < 4; ++v) EXPECT_EQ(again[v], pir[v]);
}

int main() {
  if(lpc_device_init(0) != LPC_OK) { printf("no CUDA device: %s\n", lpc_last_error()); return 2; }
  TernaryProblem();
  TemporalConstraint1Flat();
  TernaryAdds();
  MinConstraint1();
  MaxConstraint2();
  IntTimes1();
  IntDivs();
  InfiniteDomains();
  Strict1();
  InterpretationErrors();
  ExhaustiveAddWithSnapshots();
  IncrementalTells();
  Deinterpret();
  printf("%d checks, %d failures\n", g_checks, g_fail);
  return g_fail ? 1 : 0;
}
