// devhost.cu — TEST INFRASTRUCTURE: runs the device propagator code (lala-pc_b200/csrc/pir_device.cuh) on the HOST
// so that the register-level rules (incl. the flattened den_fdiv / den_cdiv) can be compared with the oracle on a
// machine without a GPU. Built by tests/test_devhost.py with nvcc; never part of the product.
#define LPC_HOST_HARNESS
#include "../../lala-pc_b200/csrc/pir_div.cuh"
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
}
