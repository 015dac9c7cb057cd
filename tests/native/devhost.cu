// devhost.cu — TEST INFRASTRUCTURE: runs the device propagator code (lala-pc_b200/csrc/pir_device.cuh) on the HOST
// so that the register-level rules (incl. the flattened den_fdiv / den_cdiv) can be compared with the oracle on a
// machine without a GPU. Built by tests/test_devhost.py with nvcc; never part of the product.
#define LPC_HOST_HARNESS
#include "../../lala-pc_b200/csrc/pir_div.cuh"
#include "../../lala-pc_b200/csrc/pc_device.cuh"
namespace lpc { LPC_HD void deduce_div(int op, Itv& r1, Itv& r2, Itv& r3) { deduce_div_rules(op, r1, r2, r3); } }
using namespace lpc;

static int to_dev(int sig) {
  switch(sig) { case 2: return D_ADD; case 4: return D_MUL; case 6: return D_MIN; case 7: return D_MAX; case 25: return D_TDIV;
    case 27: return D_FDIV; case 29: return D_CDIV; case 31: return D_EDIV; case 46: return D_EQ; default: return D_LEQ; }
}

extern "C" {
// One deduce step on an interleaved {lb,ub} store, with the commit semantics of the kernels. Returns bit0 changed,
// bit1 bot observed.
int devhost_deduce(int* lbub, const int* rec4) {
  int op = to_dev(rec4[0]), xi = rec4[1], yi = rec4[2], zi = rec4[3];
  Itv o1(lbub[2 * xi], lbub[2 * xi + 1]), o2(lbub[2 * yi], lbub[2 * yi + 1]), o3(lbub[2 * zi], lbub[2 * zi + 1]);
  Itv r1 = o1, r2 = o2, r3 = o3;
  int f = (r1.is_bot() | r2.is_bot() | r3.is_bot()) ? 2 : 0;
  deduce_regs<true>(op, r1, r2, r3);
  auto commit = [&](int v, const Itv& old, const Itv& nw) {
    int g = 0;
    if(nw.lb > old.lb) { if(nw.lb > lbub[2 * v]) lbub[2 * v] = nw.lb; g = 1; }
    if(nw.ub < old.ub) { if(nw.ub < lbub[2 * v + 1]) lbub[2 * v + 1] = nw.ub; g = 1; }
    if(g && nw.lb > nw.ub) g |= 2;
    return g;
  };
  f |= commit(xi, o1, r1) | commit(yi, o2, r2) | commit(zi, o3, r3);
  return f;
}
int devhost_ask(const int* lbub, const int* rec4) {
  int op = to_dev(rec4[0]), xi = rec4[1], yi = rec4[2], zi = rec4[3];
  return ask_regs(op, Itv(lbub[2 * xi], lbub[2 * xi + 1]), Itv(lbub[2 * yi], lbub[2 * yi + 1]), Itv(lbub[2 * zi], lbub[2 * zi + 1]));
}
// Gauss-Seidel fixpoint of n records with the device rules; stops at the first sweep that saw bot.
// Returns sweeps; *is_bot out.
int devhost_fixpoint(int* lbub, int nvars, const int* recs, long long n, int* is_bot) {
  int bot = 0;
  for(int v = 0; v < nvars; ++v) bot |= lbub[2 * v] > lbub[2 * v + 1];
  int sweeps = 0, changed = 1;
  while(changed && !bot && sweeps < 1000000) {
    changed = 0;
    for(long long i = 0; i < n; ++i) { int f = devhost_deduce(lbub, recs + 4 * i); changed |= f & 1; bot |= (f >> 1) & 1; }
    ++sweeps;
  }
  *is_bot = bot;
  return sweeps;
}
// All interval triples in [lo,hi]^3 for one op (same order as lpco_pir_exhaustive); out: 7 ints per case.
void devhost_exhaustive(int sig, int lo, int hi, int* out) {
  long long idx = 0;
  int rec[4] = {sig, 0, 1, 2};
  for(int xl = lo; xl <= hi; ++xl) for(int xu = xl; xu <= hi; ++xu)
  for(int yl = lo; yl <= hi; ++yl) for(int yu = yl; yu <= hi; ++yu)
  for(int zl = lo; zl <= hi; ++zl) for(int zu = zl; zu <= hi; ++zu, ++idx) {
    int s[6] = {xl, xu, yl, yu, zl, zu};
    if(sig == 46 || sig == 48) { if(s[0] < 0) s[0] = 0; if(s[1] > 1) s[1] = 1; }
    int bot = 0;
    devhost_fixpoint(s, 3, rec, 1, &bot);
    for(int k = 0; k < 6; ++k) out[idx * 7 + k] = s[k];
    out[idx * 7 + 6] = bot;
  }
}

// ---- PC: the flat propagators of pc_device.cuh on the host --------------------------------------------------------
struct HostAcc {
  typedef lpc::UItv Univ;
  int* d;
  mutable int seen_bot;
  int touched;
  Itv load(int v) const { if(d[2 * v] > d[2 * v + 1]) seen_bot = 1; return Itv(d[2 * v], d[2 * v + 1]); }
  int embed(int v, const Itv& u) {
    int f = 0;
    if(u.lb > d[2 * v]) { d[2 * v] = u.lb; f = 1; }
    if(u.ub < d[2 * v + 1]) { d[2 * v + 1] = u.ub; f = 1; }
    if(f && d[2 * v] > d[2 * v + 1]) f |= 2;
    touched |= f;
    return f;
  }
};
// props: n x 5 {kind, first_term, n_terms, rhs, bvar}; terms: m x 2 {coef, var}. Gauss-Seidel, stop at bot.
int devhost_pc_fixpoint(const int* props, long long n, const int* terms, int* lbub, int nvars, int* is_bot, int* has_changed) {
  int bot = 0, any = 0;
  for(int v = 0; v < nvars; ++v) bot |= lbub[2 * v] > lbub[2 * v + 1];
  int sweeps = 0, changed = 1;
  while(changed && !bot && sweeps < 100000) {
    changed = 0;
    for(long long i = 0; i < n; ++i) {
      const int* p = props + 5 * i;
      int4 h = make_int4(p[0] | (p[2] << 8), p[1], p[3], p[4]);
      HostAcc acc{lbub, 0, 0};
      int f = pc_deduce(acc, h, reinterpret_cast<const int2*>(terms) + p[1]);
      changed |= (f | acc.touched) & 1; bot |= ((f >> 1) & 1) | acc.seen_bot;
    }
    any |= changed;
    ++sweeps;
  }
  *is_bot = bot; *has_changed = any;
  return sweeps;
}
int devhost_pc_ask(const int* props, long long i, const int* terms, const int* lbub) {
  const int* p = props + 5 * i;
  int4 h = make_int4(p[0] | (p[2] << 8), p[1], p[3], p[4]);
  HostAcc acc{const_cast<int*>(lbub), 0, 0};
  return pc_ask(acc, h, reinterpret_cast<const int2*>(terms) + p[1]);
}

// ---- the same over NBitset<64> cells ---------------------------------------------------------------------------------
struct HostBitAcc {
  typedef lpc::UNb Univ;
  unsigned long long* d;
  mutable int seen_bot;
  int touched;
  u64 load(int v) const { if(d[v] == 0) seen_bot = 1; return d[v]; }
  int embed(int v, u64 u) {
    const u64 old = d[v];
    if(old == 0) return 2;
    const u64 nw = old & u;
    if(nw == old) return 0;
    d[v] = nw;
    touched |= 1;
    return nw == 0 ? 3 : 1;
  }
};
// linear kinds are handed to the tree interpreter exactly as the library's table builder does (pc_bits_view)
int devhost_pc_fixpoint_bits(const int* props_in, long long n, const int* terms_in, long long n_terms, unsigned long long* cells, int nvars,
                             int* is_bot, int* has_changed) {
  std::vector<int> vp; std::vector<int2> vt;
  lpc::pc_bits_view(props_in, n, reinterpret_cast<const int2*>(terms_in), n_terms, vp, vt);
  const int* props = vp.data();
  const int* terms = reinterpret_cast<const int*>(vt.data());
  int bot = 0, any = 0;
  for(int v = 0; v < nvars; ++v) bot |= cells[v] == 0;
  int sweeps = 0, changed = 1;
  while(changed && !bot && sweeps < 100000) {
    changed = 0;
    for(long long i = 0; i < n; ++i) {
      const int* p = props + 5 * i;
      int4 h = make_int4(p[0] | (p[2] << 8), p[1], p[3], p[4]);
      HostBitAcc acc{cells, 0, 0};
      int f = pc_deduce_bits(acc, h, reinterpret_cast<const int2*>(terms) + p[1]);
      changed |= (f | acc.touched) & 1; bot |= ((f >> 1) & 1) | acc.seen_bot;
    }
    any |= changed;
    ++sweeps;
  }
  *is_bot = bot; *has_changed = any;
  return sweeps;
}
int devhost_pc_ask_bits(const int* props_in, long long n, long long i, const int* terms_in, long long n_terms, const unsigned long long* cells) {
  std::vector<int> vp; std::vector<int2> vt;
  lpc::pc_bits_view(props_in, n, reinterpret_cast<const int2*>(terms_in), n_terms, vp, vt);
  const int* terms = reinterpret_cast<const int*>(vt.data());
  const int* p = vp.data() + 5 * i;
  int4 h = make_int4(p[0] | (p[2] << 8), p[1], p[3], p[4]);
  HostBitAcc acc{const_cast<unsigned long long*>(cells), 0, 0};
  return pc_ask_bits(acc, h, reinterpret_cast<const int2*>(terms) + p[1]);
}
// The table builder's rewriting of a flat linear propagator as a formula stream (bitset stores): words out, count returned.
int devhost_linear_tree_words(int kind, const int* terms, int n, int rhs, int bvar, int* out, int cap) {
  std::vector<int> w;
  lpc::pc_linear_tree_words(kind, reinterpret_cast<const int2*>(terms), n, rhs, bvar, w);
  if((int)w.size() > cap) return -1;
  for(size_t i = 0; i < w.size(); ++i) out[i] = w[i];
  return (int)w.size();
}
}
