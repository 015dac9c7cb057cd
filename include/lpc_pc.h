/* lpc_pc.h — C-ABI for the PC (propagator completion) hot path: n-ary terms and formulas over the interval store.
 *
 * The reference keeps each PC propagator as a heap-allocated variant tree (pc::Formula / pc::Term,
 * include/lala/formula.hpp, include/lala/terms.hpp) and walks it in PC::deduce(int) (include/lala/pc.hpp:671-680).
 * Here the common shapes are flattened into one 16-byte header + a run of {coef, var} terms, so that a sweep is a
 * streaming read instead of a pointer chase; every other shape keeps its tree as a prefix-encoded stream (LPC_PC_TREE). Each flat kind names the reference tree it stands for; the result of
 * `deduce` on it is bit-identical to walking that tree (tests/test_gpu_pc.py against oracle/pc_oracle.cpp).
 *
 * Same conventions as lpc.h (status codes, stores as {lb,ub} int32 pairs, one in-flight call per handle).
 */
#ifndef LPC_PC_H
#define LPC_PC_H

#include "lpc.h"

#ifdef __cplusplus
extern "C" {
#endif

enum lpc_pc_kind {
  /* Inequality(Nary<Add>(t_1..t_n), Constant rhs), t_i = Variable (coef 1) or Binary<Mul>(Constant coef, Variable):
   * sum_i coef_i * x_i <= rhs          formula.hpp:796-805, terms.hpp:465-499, 231-262 */
  LPC_PC_LIN_LE = 1,
  /* Biconditional(VariableLiteral(bvar), <the LIN_LE above>):  bvar <=> (sum <= rhs)      formula.hpp:421-427 */
  LPC_PC_REIF_LIN_LE = 2,
  /* Equality(Variable x, Variable y): terms = {x, y}                                       formula.hpp:672-681 */
  LPC_PC_EQ = 3,
  /* Equality<neg>(Variable x, Variable y) (2 terms) or (Variable x, Constant rhs) (1 term) formula.hpp:636-670 */
  LPC_PC_NEQ = 4,
  /* right-nested Disjunction of VariableLiterals l_1 \/ (l_2 \/ (... \/ l_n)); term coef +1 = x, -1 = not x
   * (bool_clause, pc.hpp:542-544)                                                          formula.hpp:346-350 */
  LPC_PC_CLAUSE = 5,
  /* Equality(Unary<Abs>(Variable x), Variable y): terms = {x, y}                           terms.hpp:104-119 */
  LPC_PC_ABS_EQ = 6,
  /* The other comparisons of a linear term with a constant, same term encoding as LIN_LE:
   * Inequality(Constant rhs, sum):      rhs <= sum                                          formula.hpp:801-804 */
  LPC_PC_LIN_GE = 7,
  /* Inequality<neg>(sum, Constant rhs): sum > rhs                                           formula.hpp:779-785 */
  LPC_PC_LIN_GT = 8,
  /* Equality(sum, Constant rhs):        sum = rhs                                           formula.hpp:676-680 */
  LPC_PC_LIN_EQ = 9,
  /* Equality(sum, Variable bvar):       sum = z, z in `bvar`                                formula.hpp:672-681 */
  LPC_PC_LIN_EQ_VAR = 10,
  /* Any other pc::Formula (formula.hpp:955-1002) over any pc::Term (terms.hpp:628-652): the tree itself, prefix encoded
   * as int32 words in the propagator's term slots (n_terms = ceil(words / 2), zero padded; rhs and bvar unused):
   *   terms     1 k = Constant | 2 v = Variable | 3 t = Neg | 4 t = Abs | 5 a b = Add | 6 a b = Sub | 7 a b = Mul
   *             | 8 n t1..tn = Nary<Add> | 9 a b = Min | 10 a b = Max
   *   formulas  20 v = VariableLiteral | 21 v = its negation | 22 l r = (l <= r) | 23 l r = (l > r) | 24 l r = (l = r)
   *             | 25 l r = (l != r) | 26 f g = and | 27 f g = or | 28 f g = equiv | 29 f g = imply | 30 f g = xor
   * Walked on the device by one thread per propagator exactly as Formula::deduce / Term::embed walk it
   * (lala-pc_b200/csrc/pc_tree.cuh); terms up to 8 levels and connectives up to 6 levels deep, deeper streams are
   * refused with LPC_ERR_UNSUPPORTED. Both store universes: over a bitset store the same walk runs on NBitset values
   * (see below). */
  LPC_PC_TREE = 11
};

typedef struct lpc_pc_prop {
  int32_t kind;        /* lpc_pc_kind */
  int32_t first_term;  /* index of the first term in the term array */
  int32_t n_terms;
  int32_t rhs;         /* constant right-hand side (LIN_LE, REIF_LIN_LE, NEQ with a constant) */
  int32_t bvar;        /* reification variable (REIF_LIN_LE) or the variable equal to the sum (LIN_EQ_VAR), else -1 */
} lpc_pc_prop;

typedef struct lpc_pc_term {
  int32_t coef;        /* coefficient (LIN kinds, non-zero) or literal sign (+1 / -1, CLAUSE); 1 otherwise. A coefficient
                          of -1 also stands for Unary<Neg>(Variable) and for the right operand of Binary<Sub>
                          (terms.hpp:87-102, 209-229): same bounds as Binary<Mul>(Constant -1, Variable). */
  int32_t var;
} lpc_pc_term;

typedef struct lpc_pc_table lpc_pc_table;

/* PC::deduce(const tell_type&) (pc.hpp:625-645): upload the flattened propagators, in the caller's order. */
int lpc_pc_table_create(const lpc_pc_prop* props, int64_t n_props, const lpc_pc_term* terms, int64_t n_terms,
                        int32_t nvars, lpc_pc_table** out);
int lpc_pc_table_destroy(lpc_pc_table* t);
/* PC::num_deductions (pc.hpp:665-667) */
int64_t lpc_pc_table_size(const lpc_pc_table* t);
int64_t lpc_pc_table_terms(const lpc_pc_table* t);

/* GaussSeidelIteration::fixpoint(n, PC::deduce(i)) (tests/pc_test.cpp:91-94) as one persistent kernel. */
int lpc_pc_fixpoint(const lpc_pc_table* t, lpc_store* s, const lpc_fixpoint_opts* o, lpc_fixpoint_result* r);
/* Same with HOST buffers. */
int lpc_pc_fixpoint_host(const lpc_pc_table* t, int32_t* lbub, const lpc_fixpoint_opts* o, lpc_fixpoint_result* r);
/* PC::deduce(int i) (pc.hpp:671-680), one launch. */
int lpc_pc_deduce_one(const lpc_pc_table* t, lpc_store* s, int64_t i, int* changed);
/* PC::ask(int i) over all i (pc.hpp:661-663; the loop of is_extractable, pc.hpp:726-738); bits may be NULL. */
int lpc_pc_ask_all(const lpc_pc_table* t, const lpc_store* s, int64_t* n_entailed, uint8_t* bits);

/* ---- bitset stores: VStore<NBitset<64, local_memory, unsigned long long>> (tests/pc_bitset_test.cpp:23-25) ----------
 * The same 8-byte cells of an lpc_store read as ONE uint64 per variable: bit 0 = "some value <= -1", bit i (1..62) =
 * value i - 1, bit 63 = "some value >= 62" (so [0, 61] is exact); meet = AND, bot = 0, top = all ones. A store handle
 * carries no domain tag: the caller picks the `_bits` entry points for stores it filled with lpc_store_write_bits.
 * Kinds with a lane-tile rule on bitsets: EQ, NEQ (by complement, formula.hpp:642-644), CLAUSE, ABS_EQ - the shapes the
 * reference pins in tests/pc_bitset_test.cpp. Every other kind - the LIN_* sums and LPC_PC_TREE - is walked as a formula
 * tree over the NBitset universe, one thread per propagator (the `_bits` calls rewrite the linear kinds of a table as
 * streams on first use): set operations (meet, join, complement, inclusion) are bitwise, every arithmetic operation goes
 * through the interval hull of its operands and back into the universe, so values beyond [-1, 62] fold into the two
 * open-ended bits. lala-core's NBitset::project is un-vendored and no reference test pins it beyond IntAbs1: this
 * reading is the one the CPU checker of the test-suite carries, PARITY UNPINNED upstream.
 * In that universe -x and (-1) * x differ (the constant -1 is the open-ended bit), so a front end that targets bitset
 * stores must keep Unary<Neg> / Binary<Sub> as trees instead of folding them into a coefficient of -1. */
int lpc_store_write_bits(lpc_store* s, int32_t first, int32_t n, const uint64_t* cells);
int lpc_store_read_bits(const lpc_store* s, int32_t first, int32_t n, uint64_t* cells);
/* VStore::embed / is_bot / is_top on bitset cells (pc.hpp:647-649, 685-692). */
int lpc_store_embed_bits(lpc_store* s, int32_t var, uint64_t cell, int* changed);
int lpc_store_is_bot_bits(const lpc_store* s, int* out);
int lpc_store_is_top_bits(const lpc_store* s, int* out);
/* NBitset(lb, ub): the cell holding the integer range [lb, ub] (empty if lb > ub). Pure host helper. */
uint64_t lpc_nbit_range(int32_t lb, int32_t ub);
int lpc_pc_fixpoint_bits(const lpc_pc_table* t, lpc_store* s, const lpc_fixpoint_opts* o, lpc_fixpoint_result* r);
int lpc_pc_fixpoint_bits_host(const lpc_pc_table* t, uint64_t* cells, const lpc_fixpoint_opts* o, lpc_fixpoint_result* r);
int lpc_pc_deduce_one_bits(const lpc_pc_table* t, lpc_store* s, int64_t i, int* changed);
int lpc_pc_ask_all_bits(const lpc_pc_table* t, const lpc_store* s, int64_t* n_entailed, uint8_t* bits);

#ifdef __cplusplus
}
#endif
#endif /* LPC_PC_H */
