// pc_facade_test.cpp — the reference's PC tests (lala-pc tests/pc_test.cpp, tests/pc_bitset_test.cpp) re-expressed against
// the b200pc façade, for the formula shapes that have a flat device kind. FlatZinc parsing is not part of the hot path,
// so each constraint is written as the TFormula the reference's parser hands to PC::interpret_tell; everything from
// interpret_tell on is the reference helpers' call sequence: interpret_tell -> deduce(tell) ->
// GaussSeidelIteration::fixpoint(deduce(i), has_changed) -> compare domains -> is_extractable / extract.
// Needs a CUDA device (run by tests/test_gpu_facade.py).
#include <cstdio>
#include <string>
#include <vector>

#include "../b200pc/pc.hpp"

using namespace b200pc;

static int g_fail = 0, g_checks = 0;
#define EXPECT_TRUE(c) do { ++g_checks; if(!(c)) { ++g_fail; printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); } } while(0)
#define EXPECT_FALSE(c) EXPECT_TRUE(!(c))
#define EXPECT_EQ(a, b) do { ++g_checks; if(!((a) == (b))) { ++g_fail; printf("FAIL %s:%d: %s == %s\n", __FILE__, __LINE__, #a, #b); } } while(0)

const AType sty = 0;
const AType pty = 1;

template <class L>
struct Model {
  using U = typename L::universe_type;
  VarEnv env{sty};
  std::vector<std::pair<std::string, U>> doms;
  std::vector<TF> cons;
  Model& var(const char* n, U d = U::top()) { env.declare(n); doms.push_back({n, d}); return *this; }
  Model& c(const TF& f) { cons.push_back(f); return *this; }
};

// create_and_interpret_and_tell<L>(...) of lala-core's abstract_testing.hpp
template <class L>
static L create_and_interpret_and_tell(Model<L>& m) {
  L pc(pty, std::make_shared<typename L::sub_type>(m.env.num_vars()));
  typename L::tell_type tell;
  for(auto& d : m.doms) { AVar v; m.env.interpret(F::var(d.first), v); tell.sub_value.push_back({v, d.second}); }
  std::string why;
  for(auto& f : m.cons) EXPECT_TRUE(pc.interpret_tell(f, m.env, tell, &why));
  pc.deduce(tell);
  return pc;
}

// interpret_must_succeed<IKind::TELL>(...)
template <class L>
static void tell_more(L& pc, Model<L>& m, const TF& f) {
  typename L::tell_type tell;
  std::string why;
  EXPECT_TRUE(pc.interpret_tell(f, m.env, tell, &why));
  pc.deduce(tell);
}

// tests/pc_test.cpp:69-82, tests/pc_bitset_test.cpp:29-42
template <class L>
static void test_extract(const L& pc, bool is_ua) {
  AbstractDeps deps;
  L copy1(pc, deps);
  EXPECT_EQ(pc.is_extractable(), is_ua);
  if(pc.is_extractable()) {
    pc.extract(copy1);
    EXPECT_EQ(pc.is_top(), copy1.is_top());
    EXPECT_EQ(pc.is_bot(), copy1.is_bot());
    for(int i = 0; i < pc.vars(); ++i) EXPECT_EQ(pc[i], copy1[i]);
  }
}

// tests/pc_test.cpp:84-105: once with the reference's own loop (deduce(i) one launch at a time) and once with the fused
// device fixpoint on a copy, which must agree.
template <class L>
static void deduce_and_test(L& pc, int num_deds, const std::vector<typename L::universe_type>& before,
                            const std::vector<typename L::universe_type>& after, bool is_ua, bool expect_changed = true) {
  EXPECT_EQ(pc.num_deductions(), num_deds);
  for(size_t i = 0; i < before.size(); ++i) EXPECT_EQ(pc[(int)i], before[i]);
  AbstractDeps deps;
  L fused(pc, deps);
  bool has_changed = false;
  GaussSeidelIteration{}.fixpoint(pc.num_deductions(), [&](size_t i) { return pc.deduce((int)i); }, has_changed);
  EXPECT_EQ(has_changed, expect_changed);
  for(size_t i = 0; i < after.size(); ++i) EXPECT_EQ(pc[(int)i], after[i]);
  test_extract(pc, is_ua);
  fixpoint_stats st = fused.fixpoint();
  EXPECT_FALSE(st.is_bot);
  EXPECT_EQ(st.has_changed, expect_changed);
  for(int i = 0; i < pc.vars(); ++i) EXPECT_EQ(fused[i], pc[i]);
}
template <class L>
static void deduce_and_test(L& pc, int num_deds, const std::vector<typename L::universe_type>& before_after, bool is_ua = false) {
  deduce_and_test(pc, num_deds, before_after, before_after, is_ua, false);
}

template <class L>
static void deduce_and_test_bot(L& pc, int num_deds, const std::vector<typename L::universe_type>& before) {
  EXPECT_EQ(pc.num_deductions(), num_deds);
  for(size_t i = 0; i < before.size(); ++i) EXPECT_EQ(pc[(int)i], before[i]);
  AbstractDeps deps;
  L fused(pc, deps);
  bool has_changed = false;
  GaussSeidelIteration{}.fixpoint(pc.num_deductions(), [&](size_t i) { return pc.deduce((int)i); }, [&]() { return pc.is_bot(); }, has_changed);
  EXPECT_TRUE(has_changed);
  EXPECT_TRUE(pc.is_bot());
  fixpoint_stats st = fused.fixpoint();
  EXPECT_TRUE(st.is_bot && st.has_changed && fused.is_bot());
}

static TF V(const char* n) { return TF::var(n); }
static TF K(int k) { return TF::z(k); }
static TF bin(const TF& a, int sig, const TF& b) { return TF::make_binary(a, sig, b); }
static TF times(int c, const char* x) { return bin(K(c), MUL, V(x)); }

// ---- tests/pc_test.cpp ----------------------------------------------------------------------------------------------
static void TemporalConstraint1() {   // pc_test.cpp:132-138: x + y <= 5
  Model<IPC> m;
  m.var("x", Itv(0, 10)).var("y", Itv(0, 10)).c(bin(bin(V("x"), ADD, V("y")), LEQ, K(5)));
  IPC ipc = create_and_interpret_and_tell(m);
  deduce_and_test(ipc, 1, {Itv(0, 10), Itv(0, 10)}, {Itv(0, 5), Itv(0, 5)}, false);
}

static void TemporalConstraints() {   // pc_test.cpp:114-210
  struct { TF c; std::vector<Itv> dom, after; bool ua, changed; } cases[] = {
    {bin(bin(V("x"), ADD, V("y")), EQ, K(5)), {Itv(0, 10), Itv(0, 10)}, {Itv(0, 5), Itv(0, 5)}, false, true},          // AddEquality
    {bin(bin(V("x"), ADD, V("y")), GT, K(5)), {Itv(0, 10), Itv(0, 10)}, {Itv(0, 10), Itv(0, 10)}, false, false},       // 2
    {bin(bin(V("x"), ADD, V("y")), GT, K(5)), {Itv(0, 3), Itv(0, 3)}, {Itv(3, 3), Itv(3, 3)}, true, true},             // 3
    {bin(bin(V("x"), ADD, V("y")), GEQ, K(5)), {Itv(0, 3), Itv(0, 3)}, {Itv(2, 3), Itv(2, 3)}, false, true},           // 4
    {bin(bin(V("x"), ADD, V("y")), EQ, K(5)), {Itv(0, 4), Itv(0, 4)}, {Itv(1, 4), Itv(1, 4)}, false, true},            // 5
    {bin(bin(V("x"), SUB, V("y")), LEQ, K(5)), {Itv(0, 10), Itv(0, 10)}, {Itv(0, 10), Itv(0, 10)}, false, false},      // 6
    {bin(bin(V("x"), SUB, V("y")), LEQ, K(-10)), {Itv(0, 10), Itv(0, 10)}, {Itv(0, 0), Itv(10, 10)}, true, true},      // 7
    {bin(bin(V("x"), SUB, V("y")), GEQ, K(5)), {Itv(0, 10), Itv(0, 10)}, {Itv(5, 10), Itv(0, 5)}, false, true},        // 8
    {bin(bin(V("x"), SUB, V("y")), LEQ, K(-5)), {Itv(0, 10), Itv(0, 10)}, {Itv(0, 5), Itv(5, 10)}, false, true}};      // 9
  for(auto& cs : cases) {
    Model<IPC> m;
    m.var("x", cs.dom[0]).var("y", cs.dom[1]).c(cs.c);
    IPC ipc = create_and_interpret_and_tell(m);
    deduce_and_test(ipc, 1, cs.dom, cs.after, cs.ua, cs.changed);
  }
  {   // pc_test.cpp:123-129: x + y = z, z <= 5
    Model<IPC> m;
    m.var("x", Itv(0, 10)).var("y", Itv(0, 10)).var("z").c(bin(V("z"), LEQ, K(5))).c(bin(bin(V("x"), ADD, V("y")), EQ, V("z")));
    IPC ipc = create_and_interpret_and_tell(m);
    deduce_and_test(ipc, 1, {Itv(0, 10), Itv(0, 10), Itv(INT_MIN, 5)}, {Itv(0, 5), Itv(0, 5), Itv(0, 5)}, false);
  }
}

static void NegationOps() {   // pc_test.cpp:303-357
  TF nx = TF::make_unary(NEG, V("x"));
  struct { Itv dom; TF c; Itv after; bool bot; } cases[] = {
    {Itv(-4, 3), bin(nx, LEQ, K(2)), Itv(-2, 3), false}, {Itv(-4, 3), bin(nx, LEQ, K(-2)), Itv(2, 3), false},
    {Itv(0, 3), bin(nx, LEQ, K(-2)), Itv(2, 3), false}, {Itv(-4, -3), bin(nx, LEQ, K(4)), Itv(-4, -3), false},
    {Itv(-4, 3), bin(nx, GEQ, K(-2)), Itv(-4, 2), false}, {Itv(-4, 3), bin(nx, GT, K(2)), Itv(-4, -3), false},
    {Itv(-4, 3), bin(nx, GEQ, K(5)), Itv(), true}};
  for(auto& cs : cases) {
    Model<IPC> m;
    m.var("x", cs.dom).c(cs.c);
    IPC ipc = create_and_interpret_and_tell(m);
    if(cs.bot) deduce_and_test_bot(ipc, 1, {cs.dom});
    else deduce_and_test(ipc, 1, {cs.dom}, {cs.after}, true, !(cs.after == cs.dom));
  }
}

static void TernarySums() {   // pc_test.cpp:222-260 as one n-ary sum x + y + z <= k (the shape config 3 uses)
  struct { Itv d; int k; bool bot; Itv after; bool ua; } cases[] = {
    {Itv(3, 10), 8, true, Itv(), false}, {Itv(3, 10), 9, false, Itv(3, 3), true},
    {Itv(3, 10), 10, false, Itv(3, 4), false}, {Itv(-2, 2), -5, false, Itv(-2, -1), false}};
  for(auto& cs : cases) {
    Model<IPC> m;
    m.var("x", cs.d).var("y", cs.d).var("z", cs.d).c(bin(TF::make_nary(ADD, {V("x"), V("y"), V("z")}), LEQ, K(cs.k)));
    IPC ipc = create_and_interpret_and_tell(m);
    if(cs.bot) deduce_and_test_bot(ipc, 1, {cs.d, cs.d, cs.d});
    else deduce_and_test(ipc, 1, {cs.d, cs.d, cs.d}, {cs.after, cs.after, cs.after}, cs.ua);
  }
}

static void PseudoBoolean() {   // pc_test.cpp:263-290: 2x + y + 3z <= 2 and friends, as n-ary sums
  Itv B(0, 1), Z(0, 0);
  struct { int a, b, c; std::vector<Itv> after; bool ua; } cases[] = {
    {2, 1, 3, {B, B, Z}, false}, {2, 5, 3, {B, Z, Z}, true}, {3, 5, 3, {Z, Z, Z}, true}};
  for(auto& cs : cases) {
    Model<IPC> m;
    TF y = cs.b == 1 ? V("y") : times(cs.b, "y");
    m.var("x", B).var("y", B).var("z", B).c(bin(TF::make_nary(ADD, {times(cs.a, "x"), y, times(cs.c, "z")}), LEQ, K(2)));
    IPC ipc = create_and_interpret_and_tell(m);
    deduce_and_test(ipc, 1, {B, B, B}, cs.after, cs.ua);
  }
}

static void NotEqual() {   // pc_test.cpp:386-396
  {
    Model<IPC> m;
    m.var("x", Itv(1, 10)).c(bin(V("x"), NEQ, K(10)));
    IPC ipc = create_and_interpret_and_tell(m);
    deduce_and_test(ipc, 1, {Itv(1, 10)}, {Itv(1, 9)}, true);
  }
  {
    Model<IPC> m;
    m.var("x", Itv(1, 10)).var("y", Itv(10, 10)).c(bin(V("x"), NEQ, V("y")));
    IPC ipc = create_and_interpret_and_tell(m);
    deduce_and_test(ipc, 1, {Itv(1, 10), Itv(10, 10)}, {Itv(1, 9), Itv(10, 10)}, true);
  }
}

template <class L>
static TF bool_clause() {   // bool_clause([x1, x2], [y1, y2]) == x1 \/ x2 \/ not y1 \/ not y2
  return TF::make_nary(OR, {V("x1"), V("x2"), TF::make_unary(NOT, V("y1")), TF::make_unary(NOT, V("y2"))});
}

template <class L>
static void BooleanClauses() {   // pc_test.cpp:564-610, pc_bitset_test.cpp:102-148
  using U = typename L::universe_type;
  U B(0, 1), T(1, 1), Fv(0, 0);
  auto fresh = [&](Model<L>& m) { m.var("x1", B).var("x2", B).var("y1", B).var("y2", B).c(bool_clause<L>()); return create_and_interpret_and_tell(m); };
  {
    Model<L> m; L pc = fresh(m);
    deduce_and_test(pc, 1, {B, B, B, B}, false);
    tell_more(pc, m, bin(V("x1"), EQ, K(1)));
    deduce_and_test(pc, 1, {T, B, B, B}, true);
  }
  {
    Model<L> m; L pc = fresh(m);
    deduce_and_test(pc, 1, {B, B, B, B}, false);
    tell_more(pc, m, bin(V("y1"), EQ, K(0)));
    deduce_and_test(pc, 1, {B, B, Fv, B}, true);
  }
  {
    Model<L> m; L pc = fresh(m);
    deduce_and_test(pc, 1, {B, B, B, B}, false);
    tell_more(pc, m, bin(V("x1"), EQ, K(0)));
    deduce_and_test(pc, 1, {Fv, B, B, B}, false);
    tell_more(pc, m, bin(V("x2"), EQ, K(0)));
    deduce_and_test(pc, 1, {Fv, Fv, B, B}, false);
    tell_more(pc, m, bin(V("y1"), EQ, K(1)));
    deduce_and_test(pc, 1, {Fv, Fv, T, B}, {Fv, Fv, T, Fv}, true);
  }
  {
    Model<L> m; L pc = fresh(m);
    deduce_and_test(pc, 1, {B, B, B, B}, false);
    tell_more(pc, m, bin(V("x1"), EQ, K(0)));
    deduce_and_test(pc, 1, {Fv, B, B, B}, false);
    tell_more(pc, m, bin(V("y1"), EQ, K(1)));
    deduce_and_test(pc, 1, {Fv, B, T, B}, false);
    tell_more(pc, m, bin(V("y2"), EQ, K(1)));
    deduce_and_test(pc, 1, {Fv, B, T, T}, {Fv, T, T, T}, true);
  }
}

static void IntAbs1() {   // pc_test.cpp:700-707
  Model<IPC> m;
  m.var("x", Itv(-15, 5)).var("y", Itv(-10, 10)).c(bin(TF::make_unary(ABS, V("x")), EQ, V("y")));
  IPC ipc = create_and_interpret_and_tell(m);
  deduce_and_test(ipc, 1, {Itv(-15, 5), Itv(-10, 10)}, {Itv(-10, 5), Itv(0, 10)}, false);
}

static void InfiniteDomains() {   // pc_test.cpp:766-787: b = (x <= 5)
  for(int bv = 1; bv >= 0; --bv) {
    Model<IPC> m;
    m.var("x").var("b", Itv(0, 1)).c(bin(V("b"), EQ, bin(V("x"), LEQ, K(5))));
    IPC ipc = create_and_interpret_and_tell(m);
    deduce_and_test(ipc, 1, {Itv::top(), Itv(0, 1)}, false);
    tell_more(ipc, m, bin(V("b"), EQ, K(bv)));
    if(bv) deduce_and_test(ipc, 1, {Itv::top(), Itv(1, 1)}, {Itv(INT_MIN, 5), Itv(1, 1)}, true);
    else deduce_and_test(ipc, 1, {Itv::top(), Itv(0, 0)}, {Itv(6, INT_MAX), Itv(0, 0)}, true);
  }
}

static void UnsupportedShapesAreRefused() {
  Model<IPC> m;
  m.var("x").var("y").var("z");
  IPC ipc(pty, std::make_shared<VStore>(3));
  IPC::tell_type tell;
  std::string why;
  // shapes without a flat kind keep their tree (LPC_PC_TREE) ...
  EXPECT_TRUE(ipc.interpret_tell(bin(V("x"), LEQ, V("y")), m.env, tell, &why));                       // var <= var
  EXPECT_TRUE(ipc.interpret_tell(bin(bin(V("x"), MUL, V("y")), LEQ, K(3)), m.env, tell, &why));        // non-linear
  EXPECT_TRUE(ipc.interpret_tell(bin(V("x"), AND, V("y")), m.env, tell, &why));                       // conjunction
  EXPECT_EQ((int)tell.props.size(), 3);
  for(auto& p : tell.props) EXPECT_EQ(p.kind, (int)LPC_PC_TREE);
  // ... unless they are deeper than the device interpreter walks, or hold a node PC has no rule for
  TF deep = V("x");
  for(int i = 0; i < 9; ++i) deep = bin(deep, ADD, V("y"));
  EXPECT_FALSE(ipc.interpret_tell(bin(deep, LEQ, K(3)), m.env, tell, &why));
  EXPECT_EQ(why, std::string("The shape of this formula is not supported."));
  EXPECT_FALSE(ipc.interpret_tell(TF::in(V("x"), {1, 3}), m.env, tell, &why));                         // `in` needs a bitset store
  EXPECT_EQ((int)tell.props.size(), 3);
  // sums over a bitset store are walked as trees over the NBitset universe (arithmetic through the interval hull)
  Model<BitPC> mb;
  mb.var("x", NBit(0, 5)).var("y", NBit::from_set({0, 2, 5})).c(bin(bin(V("x"), ADD, V("y")), LEQ, K(3)));
  BitPC bpc = create_and_interpret_and_tell(mb);
  deduce_and_test(bpc, 1, {NBit(0, 5), NBit::from_set({0, 2, 5})}, {NBit(0, 3), NBit::from_set({0, 2})}, false);
}

// The goldens whose formulas have no flat kind: the tree goes to the device as it is (LPC_PC_TREE).
static void TreePropagators() {
  {   // TemporalConstraint10, pc_test.cpp:213-219: x <= -5 + y
    Model<IPC> m;
    m.var("x", Itv(0, 10)).var("y", Itv(0, 10)).c(bin(V("x"), LEQ, bin(K(-5), ADD, V("y"))));
    IPC ipc = create_and_interpret_and_tell(m);
    deduce_and_test(ipc, 1, {Itv(0, 10), Itv(0, 10)}, {Itv(0, 5), Itv(5, 10)}, false);
  }
  {   // TernaryAdd2, pc_test.cpp:233-240: (x + y) + z <= 9 as nested binary sums
    Model<IPC> m;
    Itv d(3, 10);
    m.var("x", d).var("y", d).var("z", d).c(bin(bin(bin(V("x"), ADD, V("y")), ADD, V("z")), LEQ, K(9)));
    IPC ipc = create_and_interpret_and_tell(m);
    deduce_and_test(ipc, 1, {d, d, d}, {Itv(3, 3), Itv(3, 3), Itv(3, 3)}, true);
  }
  {   // ResourceConstraint1, pc_test.cpp:360-371: b <=> (x - y <= 0 /\ y - x <= 2)
    Model<IPC> m;
    m.var("x", Itv(5, 10)).var("y", Itv(9, 15)).var("b", Itv(0, 1))
     .c(bin(V("b"), EQUIV, bin(bin(bin(V("x"), SUB, V("y")), LEQ, K(0)), AND, bin(bin(V("y"), SUB, V("x")), LEQ, K(2)))));
    IPC ipc = create_and_interpret_and_tell(m);
    deduce_and_test(ipc, 1, {Itv(5, 10), Itv(9, 15), Itv(0, 1)}, false);
    tell_more(ipc, m, bin(V("b"), EQ, K(1)));
    deduce_and_test(ipc, 1, {Itv(5, 10), Itv(9, 15), Itv(1, 1)}, {Itv(7, 10), Itv(9, 12), Itv(1, 1)}, false);
  }
  {   // XorConstraint2, pc_test.cpp:438-447: x = 5 xor y = 5
    Model<IPC> m;
    m.var("x", Itv(1, 5)).var("y", Itv(1, 5)).c(bin(bin(V("x"), EQ, K(5)), XOR, bin(V("y"), EQ, K(5))));
    IPC ipc = create_and_interpret_and_tell(m);
    deduce_and_test(ipc, 1, {Itv(1, 5), Itv(1, 5)}, false);
    tell_more(ipc, m, bin(V("y"), EQ, K(5)));
    deduce_and_test(ipc, 1, {Itv(1, 5), Itv(5, 5)}, {Itv(1, 4), Itv(5, 5)}, true);
  }
  {   // MinConstraint2, pc_test.cpp:495-507: min(x, y) = z
    Model<IPC> m;
    m.var("x", Itv(0, 4)).var("y", Itv(2, 5)).var("z", Itv(0, 10)).c(bin(bin(V("x"), MIN, V("y")), EQ, V("z")));
    IPC ipc = create_and_interpret_and_tell(m);
    deduce_and_test(ipc, 1, {Itv(0, 4), Itv(2, 5), Itv(0, 10)}, {Itv(0, 4), Itv(2, 5), Itv(0, 4)}, false);
    tell_more(ipc, m, bin(V("z"), LEQ, K(3)));
    deduce_and_test(ipc, 1, {Itv(0, 4), Itv(2, 5), Itv(0, 3)}, false);
    tell_more(ipc, m, bin(V("x"), EQ, K(4)));
    deduce_and_test(ipc, 1, {Itv(4, 4), Itv(2, 5), Itv(0, 3)}, {Itv(4, 4), Itv(2, 3), Itv(2, 3)}, false);
  }
  {   // MaxConstraint2, pc_test.cpp:540-549: max(x, y) = z
    Model<IPC> m;
    m.var("x", Itv(0, 4)).var("y", Itv(2, 5)).var("z", Itv(0, 10)).c(bin(bin(V("x"), MAX, V("y")), EQ, V("z")));
    IPC ipc = create_and_interpret_and_tell(m);
    deduce_and_test(ipc, 1, {Itv(0, 4), Itv(2, 5), Itv(0, 10)}, {Itv(0, 4), Itv(2, 5), Itv(2, 5)}, false);
    tell_more(ipc, m, bin(V("z"), GEQ, K(5)));
    deduce_and_test(ipc, 1, {Itv(0, 4), Itv(2, 5), Itv(5, 5)}, {Itv(0, 4), Itv(5, 5), Itv(5, 5)}, true);
  }
  {   // IntTimes2 / IntTimes5, pc_test.cpp:626-636, 658-666: x * y = z
    Model<IPC> m;
    Itv B(0, 1);
    m.var("x", B).var("y", B).var("z", B).c(bin(bin(V("x"), MUL, V("y")), EQ, V("z")));
    IPC ipc = create_and_interpret_and_tell(m);
    deduce_and_test(ipc, 1, {B, B, B}, false);
    tell_more(ipc, m, bin(V("z"), EQ, K(1)));
    deduce_and_test(ipc, 1, {B, B, Itv(1, 1)}, {Itv(1, 1), Itv(1, 1), Itv(1, 1)}, true);
    {   // IntDiv1, pc_test.cpp:678-688: x / y = z
      Model<IPC> md;
      md.var("x", B).var("y", B).var("z", B).c(bin(bin(V("x"), TDIV, V("y")), EQ, V("z")));
      IPC ipd = create_and_interpret_and_tell(md);
      deduce_and_test(ipd, 1, {B, B, B}, {B, Itv(1, 1), B}, false);
      tell_more(ipd, md, bin(V("x"), EQ, K(1)));
      deduce_and_test(ipd, 1, {Itv(1, 1), Itv(1, 1), B}, {Itv(1, 1), Itv(1, 1), Itv(1, 1)}, true);
    }
    Model<IPC> m5;
    m5.var("x", Itv(1, 2)).var("y", B).var("z", Itv(0, 0)).c(bin(bin(V("x"), MUL, V("y")), EQ, V("z")));
    IPC ipc5 = create_and_interpret_and_tell(m5);
    deduce_and_test(ipc5, 1, {Itv(1, 2), B, Itv(0, 0)}, {Itv(1, 2), Itv(0, 0), Itv(0, 0)}, true);
  }
}

static void SnapshotRestore() {   // pc.hpp:711-723
  Model<IPC> m;
  m.var("x", Itv(0, 10)).var("y", Itv(0, 10)).c(bin(bin(V("x"), ADD, V("y")), LEQ, K(5)));
  IPC ipc = create_and_interpret_and_tell(m);
  auto snap = ipc.snapshot();
  tell_more(ipc, m, bin(V("x"), NEQ, V("y")));
  EXPECT_EQ(ipc.num_deductions(), 2);
  ipc.fixpoint();
  EXPECT_EQ(ipc[0], Itv(0, 5));
  ipc.restore(snap);
  EXPECT_EQ(ipc.num_deductions(), 1);
  EXPECT_EQ(ipc[0], Itv(0, 10));
  IPC::ask_type q;
  std::string why;
  EXPECT_TRUE(ipc.interpret_ask(bin(bin(V("x"), ADD, V("y")), LEQ, K(20)), m.env, q, &why));
  EXPECT_TRUE(ipc.ask(q));
  IPC::ask_type q2;
  EXPECT_TRUE(ipc.interpret_ask(bin(bin(V("x"), ADD, V("y")), LEQ, K(19)), m.env, q2, &why));
  EXPECT_FALSE(ipc.ask(q2));
}

// ---- tests/pc_bitset_test.cpp ---------------------------------------------------------------------------------------
static void BitNotEqual() {   // pc_bitset_test.cpp:68-90
  {
    Model<BitPC> m;
    m.var("x", NBit(1, 10)).c(bin(V("x"), NEQ, K(10)));
    BitPC bpc = create_and_interpret_and_tell(m);
    deduce_and_test(bpc, 0, {NBit(1, 9)}, {NBit(1, 9)}, true, false);
  }
  {
    Model<BitPC> m;
    m.var("x", NBit(1, 10)).var("y", NBit(10, 10)).c(bin(V("x"), NEQ, V("y")));
    BitPC bpc = create_and_interpret_and_tell(m);
    deduce_and_test(bpc, 1, {NBit(1, 10), NBit(10, 10)}, {NBit(1, 9), NBit(10, 10)}, true);
  }
  {
    Model<BitPC> m;
    m.var("x", NBit(1, 10)).c(bin(V("x"), NEQ, K(4)));
    BitPC bpc = create_and_interpret_and_tell(m);
    deduce_and_test(bpc, 0, {NBit::from_set({1, 2, 3, 5, 6, 7, 8, 9, 10})}, {NBit::from_set({1, 2, 3, 5, 6, 7, 8, 9, 10})}, true, false);
  }
  {
    Model<BitPC> m;
    m.var("x", NBit(1, 10)).var("y", NBit(4, 4)).c(bin(V("x"), NEQ, V("y")));
    BitPC bpc = create_and_interpret_and_tell(m);
    deduce_and_test(bpc, 1, {NBit(1, 10), NBit(4, 4)}, {NBit::from_set({1, 2, 3, 5, 6, 7, 8, 9, 10}), NBit(4)}, true);
  }
}

static void BitInConstraint1() {   // pc_bitset_test.cpp:93-100
  Model<BitPC> m;
  m.var("x").var("y", NBit(2, 3)).c(TF::in(V("x"), {1, 3}));
  BitPC bpc = create_and_interpret_and_tell(m);
  deduce_and_test(bpc, 0, {NBit::from_set({1, 3}), NBit(2, 3)}, true);
  tell_more(bpc, m, bin(V("x"), EQ, V("y")));
  deduce_and_test(bpc, 1, {NBit::from_set({1, 3}), NBit(2, 3)}, {NBit(3), NBit(3)}, true);
}

static void BitIntAbs1() {   // pc_bitset_test.cpp:150-157
  Model<BitPC> m;
  m.var("x", NBit(-15, 5)).var("y", NBit(-10, 10)).c(bin(TF::make_unary(ABS, V("x")), EQ, V("y")));
  BitPC bpc = create_and_interpret_and_tell(m);
  deduce_and_test(bpc, 1, {NBit(-1, 5), NBit(-1, 10)}, {NBit(-1, 5), NBit(0, 10)}, false);
}

// pc.hpp:754-788: the element back as a conjunction, with and without the entailed propagators
static size_t count_nodes(const TF& f) { size_t n = 1; for(auto& a : f.args) n += count_nodes(a); return n; }
static void Deinterpret() {
  Model<IPC> m;
  m.var("x", Itv(0, 10)).var("y", Itv(0, 10)).var("b", Itv(0, 1))
   .c(bin(bin(V("x"), ADD, V("y")), LEQ, K(5))).c(bin(V("x"), NEQ, V("y"))).c(bin(V("b"), EQUIV, bin(V("x"), LEQ, K(20))));
  IPC ipc = create_and_interpret_and_tell(m);
  TF f = ipc.deinterpret(m.env);
  EXPECT_EQ(f.sig(), (int)AND);
  EXPECT_EQ((int)f.args.size(), 1 + 3);                 // the store, then one formula per propagator
  EXPECT_EQ((int)f.seq(0).args.size(), 6);              // two bounds for each of x, y, b
  ipc.fixpoint();
  // b <=> (x <= 20) is entailed once b = 1 (x <= 5 <= 20); x + y <= 5 and x != y are not
  size_t ent = 0;
  TF g = ipc.deinterpret(m.env, true, ent);
  EXPECT_EQ(ent, (size_t)1);
  EXPECT_EQ((int)g.args.size(), 1 + 2);
  EXPECT_EQ(ipc[2], Itv(1, 1));
  // telling the deinterpreted formulas to a fresh element gives the same element
  Model<IPC> m2;
  m2.var("x").var("y").var("b");
  IPC again(pty, std::make_shared<VStore>(3));
  IPC::tell_type t2;
  std::string why;
  TF all = ipc.deinterpret(m.env);
  for(auto& a : all.seq(0).args) EXPECT_TRUE(again.interpret_tell(a, m2.env, t2, &why));
  for(size_t i = 1; i < all.args.size(); ++i) EXPECT_TRUE(again.interpret_tell(all.args[i], m2.env, t2, &why));
  again.deduce(t2);
  EXPECT_EQ(again.num_deductions(), ipc.num_deductions());
  for(int v = 0; v < 3; ++v) EXPECT_EQ(again[v], ipc[v]);
  EXPECT_TRUE(count_nodes(all) > 10);
  // without propagators deinterpret is the store alone (pc.hpp:757-759)
  IPC empty(pty, std::make_shared<VStore>(3));
  EXPECT_EQ((int)empty.deinterpret(m2.env).args.size(), 0);
  // a bitset store with a hole deinterprets to `x in S`
  Model<BitPC> mb;
  mb.var("x", NBit(1, 6)).c(TF::in(V("x"), {1, 2, 4, 6}));
  BitPC bpc = create_and_interpret_and_tell(mb);
  TF fb = bpc.deinterpret(mb.env);
  EXPECT_EQ(fb.sig(), (int)AND);
  EXPECT_EQ((int)fb.args.size(), 1);
  EXPECT_EQ(fb.seq(0).sig(), (int)IN);
  EXPECT_EQ((int)fb.seq(0).seq(1).set.size(), 4);
}

int main() {
  if(lpc_device_init(0) != LPC_OK) { printf("no CUDA device: %s\n", lpc_last_error()); return 2; }
  TemporalConstraint1();
  TemporalConstraints();
  NegationOps();
  TernarySums();
  PseudoBoolean();
  NotEqual();
  BooleanClauses<IPC>();
  IntAbs1();
  InfiniteDomains();
  UnsupportedShapesAreRefused();
  TreePropagators();
  SnapshotRestore();
  Deinterpret();
  BitNotEqual();
  BitInConstraint1();
  BooleanClauses<BitPC>();
  BitIntAbs1();
  printf("%d checks, %d failures\n", g_checks, g_fail);
  return g_fail ? 1 : 0;
}
