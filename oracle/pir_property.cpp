// pir_property.cpp — TEST INFRASTRUCTURE ONLY (see pir_oracle.cpp).
//
// Restatement of the reference's exhaustive bounds-consistency property harness
// (tests/bound_consistency_test.hpp:155-225, driven by tests/pir_test.cpp:122-137) applied to the oracle:
// for every interval triple in [minval,maxval]^3 and one ternary record `x = y op z`,
//   - brute-force the concrete solutions and their hull (:183-193),
//   - run the fixpoint (:36-39),
//   - bot is only allowed when there is no solution (:40-42),
//   - the result must equal the hull (test_completeness) or contain it (:44-51),
//   - if every propagator is `ask`-entailed the hull must be non-empty and every point of the hull box must
//     satisfy the predicate (:198-217).
// It also exports the fixpoints themselves so that the GPU batched kernel can be compared bit-for-bit on the
// same 231^3 stores.

#include <cstdint>
#include <climits>
#include <vector>
#include <thread>
#include <atomic>
#include <algorithm>

extern "C" {
struct lpco_stats { int32_t has_changed, is_bot; int64_t sweeps, deductions; double seconds; };
void lpco_pir_fixpoint(int32_t* lbub, int32_t nvars, const int32_t* recs, int64_t n, int32_t stop_on_bot,
                       int64_t max_sweeps, lpco_stats* out);
int lpco_pir_ask(const int32_t* lbub, int32_t nvars, const int32_t* rec4);
int32_t lpco_div(int32_t a, int32_t op, int32_t b);
}

namespace {
enum { ADD = 2, MUL = 4, MIN = 6, MAX = 7, TDIV = 25, FDIV = 27, CDIV = 29, EDIV = 31, EQ = 46, LEQ = 48 };

// The concrete predicates of tests/pir_test.cpp:123-132.
inline bool pred(int op, int x, int y, int z) {
  switch(op) {
    case EQ: return (x == 0 || x == 1) && x == (y == z);
    case LEQ: return (x == 0 || x == 1) && x == (y <= z);
    case ADD: return x == y + z;
    case MIN: return x == std::min(y, z);
    case MAX: return x == std::max(y, z);
    case MUL: return x == y * z;
    default: return z != 0 && x == lpco_div(y, op, z);
  }
}

struct Counters {
  std::atomic<int64_t> cases{0}, bot_cases{0}, unsound{0}, incomplete{0}, spurious_bot{0}, bad_entail{0},
    entailed_cases{0}, not_converged{0};
};

void run_slice(int op, int minval, int maxval, int complete, int xl, Counters& c, int32_t* out_fix, int64_t base) {
  const int32_t rec[4] = {op, 0, 1, 2};
  int64_t idx = base;
  for(int xu = xl; xu <= maxval; ++xu)
  for(int yl = minval; yl <= maxval; ++yl)
  for(int yu = yl; yu <= maxval; ++yu)
  for(int zl = minval; zl <= maxval; ++zl)
  for(int zu = zl; zu <= maxval; ++zu, ++idx) {
    c.cases++;
    int32_t s[6] = {xl, xu, yl, yu, zl, zu};
    if(op == EQ || op == LEQ) { s[0] = std::max(s[0], 0); s[1] = std::min(s[1], 1); }  // pir.hpp:333-335
    // hull of the concrete solutions
    int hx0 = INT_MAX, hx1 = INT_MIN, hy0 = INT_MAX, hy1 = INT_MIN, hz0 = INT_MAX, hz1 = INT_MIN;
    for(int a = xl; a <= xu; ++a) for(int b = yl; b <= yu; ++b) for(int d = zl; d <= zu; ++d)
      if(pred(op, a, b, d)) {
        hx0 = std::min(hx0, a); hx1 = std::max(hx1, a);
        hy0 = std::min(hy0, b); hy1 = std::max(hy1, b);
        hz0 = std::min(hz0, d); hz1 = std::max(hz1, d);
      }
    bool hull_bot = hx0 > hx1;
    lpco_stats st;
    lpco_pir_fixpoint(s, 3, rec, 1, 1, 1000000, &st);
    if(st.sweeps >= 1000000) c.not_converged++;
    if(out_fix) { for(int k = 0; k < 6; ++k) out_fix[idx * 7 + k] = s[k]; out_fix[idx * 7 + 6] = st.is_bot; }
    if(st.is_bot) {
      c.bot_cases++;
      if(!hull_bot) c.spurious_bot++;
      continue;
    }
    if(complete) {
      if(hull_bot || s[0] != hx0 || s[1] != hx1 || s[2] != hy0 || s[3] != hy1 || s[4] != hz0 || s[5] != hz1) c.incomplete++;
    }
    else if(!hull_bot) {
      if(s[0] > hx0 || s[1] < hx1 || s[2] > hy0 || s[3] < hy1 || s[4] > hz0 || s[5] < hz1) c.unsound++;
    }
    bool ent = lpco_pir_ask(s, 3, rec);
    if(ent) {
      c.entailed_cases++;
      if(hull_bot) c.bad_entail++;
      else {
        bool ok = true;
        for(int a = hx0; a <= hx1 && ok; ++a) for(int b = hy0; b <= hy1 && ok; ++b) for(int d = hz0; d <= hz1 && ok; ++d)
          ok = pred(op, a, b, d);
        if(!ok) c.bad_entail++;
      }
    }
  }
}
} // namespace

extern "C" {

// Number of interval triples in [minval,maxval]^3.
int64_t lpco_pir_exhaustive_count(int minval, int maxval) {
  int64_t w = maxval - minval + 1, m = w * (w + 1) / 2;
  return m * m * m;
}

// out[8] = {cases, bot_cases, unsound, incomplete, spurious_bot, bad_entail, entailed_cases, not_converged}.
// out_fix (optional): per case 7 int32 {xl,xu,yl,yu,zl,zu,is_bot} of the fixpoint, in loop order
// (xl, xu, yl, yu, zl, zu ascending, upper >= lower).
void lpco_pir_exhaustive(int op, int minval, int maxval, int complete, int threads, int64_t* out, int32_t* out_fix) {
  Counters c;
  int64_t w = maxval - minval + 1, m = w * (w + 1) / 2;
  std::vector<int64_t> base(w + 1, 0);   // first case index of each xl slice
  for(int i = 0; i < w; ++i) base[i + 1] = base[i] + (w - i) * m * m;
  std::atomic<int> next{0};
  auto work = [&]() {
    for(;;) {
      int i = next++;
      if(i >= w) break;
      run_slice(op, minval, maxval, complete, minval + i, c, out_fix, base[i]);
    }
  };
  std::vector<std::thread> th;
  for(int t = 1; t < std::max(1, threads); ++t) th.emplace_back(work);
  work();
  for(auto& t : th) t.join();
  out[0] = c.cases; out[1] = c.bot_cases; out[2] = c.unsound; out[3] = c.incomplete; out[4] = c.spurious_bot;
  out[5] = c.bad_entail; out[6] = c.entailed_cases; out[7] = c.not_converged;
}

} // extern "C"
