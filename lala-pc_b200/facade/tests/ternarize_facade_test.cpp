// ternarize_facade_test.cpp — the FlatZinc models of lala-pc's tests/pir_test.cpp, which reach PIR through lala-core's
// ternariser there, run here through b200pc::Ternarizer -> PIR::interpret_tell -> the device fixpoint. Expected
// intervals (and, where the decomposition is determined, deduction counts) are the reference's own.
// Needs a CUDA device (run by tests/test_gpu_facade.py).
#include <cstdio>
#include <string>
#include <vector>

#include "../b200pc/ternarize.hpp"

using namespace b200pc;

static int g_fail = 0, g_checks = 0;
#define EXPECT_TRUE(c) do { ++g_checks; if(!(c)) { ++g_fail; printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); } } while(0)
#define EXPECT_EQ(a, b) do { ++g_checks; if(!((a) == (b))) { ++g_fail; printf("FAIL %s:%d: %s == %s\n", __FILE__, __LINE__, #a, #b); } } while(0)

static TF V(const char* n) { return TF::var(n); }
static TF K(int k) { return TF::z(k); }
static TF bin(const TF& a, int sig, const TF& b) { return TF::make_binary(a, sig, b); }

struct Case {
  VarEnv env{0};
  std::map<std::string, Itv> doms;
  std::vector<std::string> order;
  std::vector<TF> cons;
  Case& var(const char* n, Itv d = Itv::top()) { env.declare(n); doms[n] = d; order.push_back(n); return *this; }
  Case& c(const TF& f) { cons.push_back(f); return *this; }
  // ternarise, interpret, propagate on the device; returns the PIR (original variables first, in declaration order)
  PIR run(int expect_deds = -1) {
    Ternarizer tz(env, doms);
    for(auto& f : cons) EXPECT_TRUE(tz.tell(f));
    PIR pir(1, std::make_shared<VStore>(env.num_vars()));
    PIR::tell_type tell;
    AVar v;
    for(auto& n : order) { env.interpret(F::var(n), v); tell.sub_value.push_back({v, doms[n]}); }
    for(auto& nv : tz.new_vars()) { env.interpret(F::var(nv.first), v); tell.sub_value.push_back({v, nv.second}); }
    std::string why;
    for(auto& f : tz.constraints()) EXPECT_TRUE(pir.interpret_tell(f, env, tell, &why));
    pir.deduce(tell);
    if(expect_deds >= 0) EXPECT_EQ(pir.num_deductions(), expect_deds);
    last_stats = pir.fixpoint();
    return pir;
  }
  fixpoint_stats last_stats;
};

static void expect_vars(const PIR& pir, const std::vector<Itv>& after) {
  for(size_t i = 0; i < after.size(); ++i) EXPECT_EQ(pir[(int)i], after[i]);
}

int main() {
  if(lpc_device_init(0) != LPC_OK) { printf("no CUDA device: %s\n", lpc_last_error()); return 2; }
  Itv D(0, 10), B(0, 1);
  { Case m; m.var("x", D).var("y", D).c(bin(bin(V("x"), ADD, V("y")), LEQ, K(5)));            // TemporalConstraint1, pir_test.cpp:232-238
    expect_vars(m.run(1), {Itv(0, 5), Itv(0, 5)}); }
  { Case m; m.var("x", D).var("y", D).c(bin(bin(V("x"), ADD, V("y")), GT, K(5)));             // TemporalConstraint2, :241-247
    expect_vars(m.run(1), {D, D}); }
  { Case m; m.var("x", Itv(0, 3)).var("y", Itv(0, 3)).c(bin(bin(V("x"), ADD, V("y")), GT, K(5)));   // TemporalConstraint3, :250-256
    PIR p = m.run(1); expect_vars(p, {Itv(3, 3), Itv(3, 3)}); EXPECT_TRUE(p.is_extractable()); }
  { Case m; m.var("x", Itv(0, 3)).var("y", Itv(0, 3)).c(bin(bin(V("x"), ADD, V("y")), GEQ, K(5)));  // TemporalConstraint4, :259-265
    expect_vars(m.run(1), {Itv(2, 3), Itv(2, 3)}); }
  { Case m; m.var("x", Itv(0, 4)).var("y", Itv(0, 4)).c(bin(bin(V("x"), ADD, V("y")), EQ, K(5)));   // TemporalConstraint5, :268-274
    expect_vars(m.run(1), {Itv(1, 4), Itv(1, 4)}); }
  { Case m; m.var("x", D).var("y", D).c(bin(bin(V("x"), SUB, V("y")), LEQ, K(-10)));          // TemporalConstraint7, :286-292
    PIR p = m.run(1); expect_vars(p, {Itv(0, 0), Itv(10, 10)}); EXPECT_TRUE(p.is_extractable()); }
  { Case m; m.var("x", D).var("y", D).c(bin(bin(V("x"), SUB, V("y")), GEQ, K(5)));            // TemporalConstraint8, :295-301
    expect_vars(m.run(1), {Itv(5, 10), Itv(0, 5)}); }
  { Case m; m.var("x", D).var("y", D).c(bin(V("x"), LEQ, bin(K(-5), ADD, V("y"))));           // TemporalConstraint10, :313-319
    expect_vars(m.run(2), {Itv(0, 5), Itv(5, 10)}); }
  { Case m; Itv d(3, 10); m.var("x", d).var("y", d).var("z", d).c(bin(bin(bin(V("x"), ADD, V("y")), ADD, V("z")), LEQ, K(8)));   // TopProp, :322-329
    PIR p = m.run(2); EXPECT_TRUE(m.last_stats.is_bot && p.is_bot()); }
  { Case m; Itv d(3, 10); m.var("x", d).var("y", d).var("z", d).c(bin(bin(bin(V("x"), ADD, V("y")), ADD, V("z")), LEQ, K(9)));   // TernaryAdd2, :332-339
    PIR p = m.run(2); expect_vars(p, {Itv(3, 3), Itv(3, 3), Itv(3, 3)}); EXPECT_TRUE(p.is_extractable()); }
  { Case m; Itv d(-2, 2); m.var("x", d).var("y", d).var("z", d).c(bin(bin(bin(V("x"), ADD, V("y")), ADD, V("z")), LEQ, K(-5)));  // TernaryAdd4, :352-359
    expect_vars(m.run(2), {Itv(-2, -1), Itv(-2, -1), Itv(-2, -1)}); }
  { Case m; m.var("x", B).var("y", B).var("z", B)                                              // PseudoBoolean1, :362-369
      .c(bin(bin(bin(bin(K(2), MUL, V("x")), ADD, V("y")), ADD, bin(K(3), MUL, V("z"))), LEQ, K(2)));
    expect_vars(m.run(4), {B, B, Itv(0, 0)}); }
  { Case m; m.var("x", B).var("y", B).var("z", B)                                              // PseudoBoolean2, :372-378
      .c(bin(bin(bin(bin(K(2), MUL, V("x")), ADD, bin(K(5), MUL, V("y"))), ADD, bin(K(3), MUL, V("z"))), LEQ, K(2)));
    expect_vars(m.run(5), {B, Itv(0, 0), Itv(0, 0)}); }
  { Case m; m.var("x", Itv(1, 10)).c(bin(V("x"), NEQ, K(10)));                                  // NotEqualConstraint1, :544-548
    PIR p = m.run(1); expect_vars(p, {Itv(1, 9)}); EXPECT_TRUE(p.is_extractable()); }
  { Case m; m.var("x", Itv(1, 10)).var("y", Itv(1, 1)).c(bin(V("x"), NEQ, V("y")));             // NotEqualConstraint3, :556-560
    expect_vars(m.run(1), {Itv(2, 10), Itv(1, 1)}); }
  { Case m; m.var("x", Itv(1, 10)).c(TF::make_unary(NOT, bin(V("x"), EQ, K(10))));               // NotEqualConstraint4, :562-566
    expect_vars(m.run(1), {Itv(1, 9)}); }
  { Case m; m.var("x", Itv(1, 1)).var("y").c(bin(bin(V("x"), EQ, K(5)), XOR, bin(V("y"), EQ, K(5))));   // XorConstraint1 after x = 1, :586-594
    PIR p = m.run(3); expect_vars(p, {Itv(1, 1), Itv(5, 5)}); EXPECT_TRUE(p.is_extractable()); }
  { Case m; m.var("x", Itv(1, 5)).var("y", Itv(5, 5)).c(bin(bin(V("x"), EQ, K(5)), XOR, bin(V("y"), EQ, K(5))));   // XorConstraint2, :597-605
    expect_vars(m.run(3), {Itv(1, 4), Itv(5, 5)}); }
  { Case m; m.var("x").var("b1", Itv(1, 1)).var("b2", Itv(1, 1))                               // MinConstraint3 last step, :657-668
      .c(bin(bin(V("b1"), MIN, V("b2")), EQ, K(1))).c(bin(V("b1"), EQ, bin(V("x"), LEQ, K(5)))).c(bin(V("b2"), EQ, bin(V("x"), GEQ, K(5))));
    expect_vars(m.run(3), {Itv(5, 5), Itv(1, 1), Itv(1, 1)}); }
  { Case m; m.var("x").var("b1", Itv(0, 1)).var("b2", Itv(0, 1))                               // MaxConstraint3, :698-709
      .c(bin(bin(V("b1"), MAX, V("b2")), EQ, K(0))).c(bin(V("b1"), EQ, bin(V("x"), LEQ, K(5)))).c(bin(V("b2"), EQ, bin(V("x"), GEQ, K(7))));
    expect_vars(m.run(3), {Itv(6, 6), Itv(0, 0), Itv(0, 0)}); }
  { Case m; m.var("x1", Itv(0, 0)).var("x2", B).var("y1", Itv(1, 1)).var("y2", Itv(1, 1))      // BooleanClause4 last step, :745-757
      .c(TF::make_nary(OR, {V("x1"), V("x2"), TF::make_unary(NOT, V("y1")), TF::make_unary(NOT, V("y2"))}));
    PIR p = m.run(5); expect_vars(p, {Itv(0, 0), Itv(1, 1), Itv(1, 1), Itv(1, 1)}); EXPECT_TRUE(p.is_extractable()); }
  { Case m; m.var("x", B).var("y", B).var("z", Itv(1, 1)).c(bin(bin(V("x"), MUL, V("y")), EQ, V("z")));   // IntTimes2, :773-783
    expect_vars(m.run(1), {Itv(1, 1), Itv(1, 1), Itv(1, 1)}); }
  { Case m; m.var("x", Itv(-15, 5)).var("y", Itv(-10, 10)).c(bin(TF::make_unary(ABS, V("x")), EQ, V("y")));   // int_abs(x, y)
    expect_vars(m.run(3), {Itv(-10, 5), Itv(0, 10)}); }                                          // IntAbs1, pir_test.cpp:915-922
  { Case m; m.var("x", Itv(5, 10)).var("y", Itv(9, 15)).var("b", Itv(1, 1))                    // ResourceConstraint1 with b = 1, pir_test.cpp:476-487
      .c(bin(V("b"), EQUIV, bin(bin(bin(V("x"), SUB, V("y")), LEQ, K(0)), AND, bin(bin(V("y"), SUB, V("x")), LEQ, K(2)))));
    expect_vars(m.run(5), {Itv(7, 10), Itv(9, 12), Itv(1, 1)}); }
  printf("%d checks, %d failures\n", g_checks, g_fail);
  return g_fail ? 1 : 0;
}
