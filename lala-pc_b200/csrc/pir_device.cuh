// pir_device.cuh — the PIR ternary propagators in registers, for sm_100a.
//
// Device counterpart of PIR::deduce(bytecode_type) (lala-pc include/lala/pir.hpp:721-817) and
// PIR::ask(bytecode_type) (pir.hpp:417-438). The rules are evaluated on three (lb, ub) register pairs; the
// caller owns the store access (global memory + L2 atomics, or shared memory) through the `commit` step:
// new bounds are joined into the store with atomicMax(lb) / atomicMin(ub), issued only when they tighten.
//
// Differences from the reference text, none of which changes a result:
//  * den_fdiv / den_cdiv (pir.hpp:520-630) recurse on sign-split sub-intervals; here they are flattened into
//    "hull of the applicable parts" (negative part, {-1} or {1}, {0}, positive part).
//  * cdiv and fdiv of the same operands share one hardware division (mul_inv, pir.hpp:715-716).
//  * int32 arithmetic wraps (the reference is UB on overflow); division is total (b == 0 -> 0).
#pragma once
#include <cstdint>

namespace lpc {

// Device opcodes (dense, so that warps of an op-sorted table are uniform and the switch is a jump table).
enum DevOp : int { D_ADD = 0, D_MUL = 1, D_MIN = 2, D_MAX = 3, D_TDIV = 4, D_FDIV = 5, D_CDIV = 6, D_EDIV = 7,
                   D_EQ = 8, D_LEQ = 9, D_NOP = 255 };

#define LPC_INF 2147483647
#define LPC_MINF (-2147483647 - 1)

struct Itv {
  int lb, ub;
  __host__ __device__ __forceinline__ Itv() {}
  __host__ __device__ __forceinline__ Itv(int l, int u) : lb(l), ub(u) {}
  __host__ __device__ __forceinline__ bool is_bot() const { return lb > ub; }
  __host__ __device__ __forceinline__ void meet(const Itv& o) { lb = lb > o.lb ? lb : o.lb; ub = ub < o.ub ? ub : o.ub; }
};

__device__ __forceinline__ Itv itv_top() { return Itv(LPC_MINF, LPC_INF); }
__device__ __forceinline__ Itv itv_bot() { return Itv(LPC_INF, LPC_MINF); }
// hull ignoring empty operands (lala-core Interval::join / fjoin)
__device__ __forceinline__ Itv fjoin(const Itv& a, const Itv& b) {
  if(b.is_bot()) return a;
  if(a.is_bot()) return b;
  return Itv(min(a.lb, b.lb), max(a.ub, b.ub));
}

__device__ __forceinline__ int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
__device__ __forceinline__ int wsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
__device__ __forceinline__ int wmul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }
__device__ __forceinline__ int wneg(int a) { return (int)(0u - (unsigned)a); }

// One hardware division, both roundings. b == 0 -> {0,0}; b == -1 -> wrap negate (INT_MIN / -1 does not trap).
struct QR { int q, r; };
__device__ __forceinline__ QR divqr(int a, int b) {
  QR o;
  if(b == 0) { o.q = 0; o.r = 0; return o; }
  if(b == -1) { o.q = wneg(a); o.r = 0; return o; }
  o.q = a / b;
  o.r = a - o.q * b;
  return o;
}
__device__ __forceinline__ int fdiv_of(QR d, int b) { return (d.r != 0 && ((d.r < 0) != (b < 0))) ? d.q - 1 : d.q; }
__device__ __forceinline__ int cdiv_of(QR d, int b) { return (d.r != 0 && ((d.r < 0) == (b < 0))) ? d.q + 1 : d.q; }
__device__ __forceinline__ int ediv_of(QR d, int b) { return d.r >= 0 ? d.q : (b > 0 ? d.q - 1 : d.q + 1); }
__device__ __forceinline__ int fdiv(int a, int b) { return fdiv_of(divqr(a, b), b); }
__device__ __forceinline__ int cdiv(int a, int b) { return cdiv_of(divqr(a, b), b); }
// pir.hpp:407-415
__device__ __forceinline__ int divop(int a, int op, int b) {
  QR d = divqr(a, b);
  return op == D_TDIV ? d.q : op == D_FDIV ? fdiv_of(d, b) : op == D_CDIV ? cdiv_of(d, b) : ediv_of(d, b);
}

#define xl r1.lb
#define xu r1.ub
#define yl r2.lb
#define yu r2.ub
#define zl r3.lb
#define zu r3.ub

// pir.hpp:702-718
__device__ __forceinline__ void mul_inv(const Itv& r1, Itv& r2, Itv& r3) {
  const bool xnz = xl > 0 || xu < 0;
  if(xnz) {
    if(zl == 0) r3.lb = 1;
    if(zu == 0) r3.ub = -1;
  }
  if(xnz && zl < 0 && zu > 0) {
    r2.lb = max(yl, min(xl, xu == LPC_MINF ? LPC_INF : wneg(xu)));
    r2.ub = min(yu, max(xl == LPC_INF ? LPC_MINF : wneg(xl), xu));
  }
  else if(xnz || zl > 0 || zu < 0) {
    if(xl == LPC_MINF || xu == LPC_INF || zl == LPC_MINF || zu == LPC_INF) return;
    if(r3.is_bot()) return;
    QR a = divqr(xl, zl), c = divqr(xu, zl);
    int lo = min(cdiv_of(a, zl), cdiv_of(c, zl));
    int hi = max(fdiv_of(a, zl), fdiv_of(c, zl));
    if(zu != zl) {   // a constant factor needs two divisions instead of four
      QR b = divqr(xl, zu), d = divqr(xu, zu);
      lo = min(lo, min(cdiv_of(b, zu), cdiv_of(d, zu)));
      hi = max(hi, max(fdiv_of(b, zu), fdiv_of(d, zu)));
    }
    r2.lb = max(yl, lo);
    r2.ub = min(yu, hi);
  }
}

// pir.hpp:449-467 — r1 = r2 / r3
__device__ __forceinline__ void itv_div(int op, Itv& r1, Itv& r2, Itv& r3) {
  if(zl < 0 && zu > 0) {
    r1.lb = max(xl, min(yl, yu == LPC_MINF ? LPC_INF : wneg(yu)));
    r1.ub = min(xu, max(yl == LPC_INF ? LPC_MINF : wneg(yl), yu));
  }
  else {
    if(zl == 0) r3.lb = 1;
    if(zu == 0) r3.ub = -1;
    if(yl == LPC_MINF || yu == LPC_INF || zl == LPC_MINF || zu == LPC_INF) return;
    if(r3.is_bot()) return;
    int t1 = divop(yl, op, zl), t2 = divop(yl, op, zu), t3 = divop(yu, op, zl), t4 = divop(yu, op, zu);
    r1.lb = max(xl, min(min(t1, t2), min(t3, t4)));
    r1.ub = min(xu, max(max(t1, t2), max(t3, t4)));
  }
}

// pir.hpp:469-479
__device__ __forceinline__ Itv num_fdiv(const Itv& r1, const Itv& r3) {
  const int xu1 = wadd(xu, 1);
  if(zl < 0 && zu > 0) {
    return Itv(min(min(xl, wneg(xu)), min(wmul(xl, zu), wadd(wmul(xu1, zl), 1))),
               max(max(wneg(xl), xu), max(wmul(xl, zl), wsub(wmul(xu1, zu), 1))));
  }
  else if(zl > 0 || zu < 0) {
    return Itv(min(min(wmul(xl, zl), wmul(xl, zu)), min(wadd(wmul(xu1, zl), 1), wadd(wmul(xu1, zu), 1))),
               max(max(wmul(xl, zl), wmul(xl, zu)), max(wsub(wmul(xu1, zl), 1), wsub(wmul(xu1, zu), 1))));
  }
  return itv_top();
}

// pir.hpp:481-491
__device__ __forceinline__ Itv num_cdiv(const Itv& r1, const Itv& r3) {
  const int xl1 = wsub(xl, 1);
  if(zl < 0 && zu > 0) {
    return Itv(min(min(xl, wneg(xu)), min(wmul(xu, zl), wadd(wmul(xl1, zu), 1))),
               max(max(wneg(xl), xu), max(wmul(xu, zu), wsub(wmul(xl1, zl), 1))));
  }
  else if(zl > 0 || zu < 0) {
    return Itv(min(min(wmul(xu, zl), wmul(xu, zu)), min(wadd(wmul(xl1, zl), 1), wadd(wmul(xl1, zu), 1))),
               max(max(wmul(xu, zl), wmul(xu, zu)), max(wsub(wmul(xl1, zl), 1), wsub(wmul(xl1, zu), 1))));
  }
  return itv_top();
}

// pir.hpp:493-507
__device__ __forceinline__ Itv num_tdiv(const Itv& r1, const Itv& r3) {
  if(xl > 0) return num_fdiv(r1, r3);
  else if(xu < 0) return num_cdiv(r1, r3);
  else if(xl <= 0 && 0 <= xu) {
    Itv r(wadd(min(zl, wneg(zu)), 1), wsub(max(wneg(zl), zu), 1));
    if(xl != 0) r = fjoin(r, num_cdiv(Itv(xl, -1), r3));
    if(xu != 0) r = fjoin(r, num_fdiv(Itv(1, xu), r3));
    return r;
  }
  return itv_top();
}

// pir.hpp:510-517
__device__ __forceinline__ Itv num_ediv(const Itv& r1, const Itv& r3) {
  if(zl > 0) return num_fdiv(r1, r3);
  else if(zu < 0) return num_cdiv(r1, r3);
  else if(zl < 0 && zu > 0) return fjoin(num_cdiv(r1, Itv(zl, -1)), num_fdiv(r1, Itv(1, zu)));
  return itv_top();
}

// ---- den_fdiv (pir.hpp:520-574), flattened ----------------------------------------------------------------------
// Branch `xl > 0 || xu + 1 < 0` for y > 0 / y < 0 (pir.hpp:523-534).
__device__ __forceinline__ Itv den_fdiv_pos(const Itv& r1, int ylo, int yhi) {   // y in [ylo,yhi], ylo > 0
  const int xu1 = wadd(xu, 1);
  return Itv(wadd(min(fdiv(ylo, xu1), fdiv(yhi, xu1)), 1), max(fdiv(ylo, xl), fdiv(yhi, xl)));
}
__device__ __forceinline__ Itv den_fdiv_neg(const Itv& r1, int ylo, int yhi) {   // yhi < 0
  const int xu1 = wadd(xu, 1);
  return Itv(min(cdiv(ylo, xl), cdiv(yhi, xl)), wsub(max(cdiv(ylo, xu1), cdiv(yhi, xu1)), 1));
}
// x definitely outside {-1, 0}: pir.hpp:522-547
__device__ __forceinline__ Itv den_fdiv_A(const Itv& r1, const Itv& r2) {
  if(yl > 0) return den_fdiv_pos(r1, yl, yu);
  if(yu < 0) return den_fdiv_neg(r1, yl, yu);
  if(yl == 0 && yu == 0) return itv_bot();
  Itv r = itv_bot();
  if(yl < 0) r = den_fdiv_neg(r1, yl, -1);
  if(yu > 0) r = fjoin(r, den_fdiv_pos(r1, 1, yu));
  return r;
}
// x = [0,0]: pir.hpp:548-552
__device__ __forceinline__ Itv den_fdiv_B(const Itv& r2) {
  if(yl > 0) return Itv(wadd(yl, 1), LPC_INF);
  if(yu < 0) return Itv(LPC_MINF, wsub(yu, 1));
  return itv_top();
}
// x = [-1,-1]: pir.hpp:553-559
__device__ __forceinline__ Itv den_fdiv_C(const Itv& r2) {
  if(yl > 0) return Itv(LPC_MINF, wneg(yl));
  if(yu < 0) return Itv(wneg(yu), LPC_INF);
  if(0 == yl && yl < yu) return Itv(LPC_MINF, -1);
  if(yl < yu && yu == 0) return Itv(1, LPC_INF);
  if(yl == 0 && yu == 0) return itv_bot();
  return itv_top();
}
static __device__ __noinline__ Itv den_fdiv(const Itv& r1, const Itv& r2) {
  if(xl > 0 || xu < -1) return den_fdiv_A(r1, r2);
  if(xl > xu) return itv_top();
  Itv r = itv_bot();
  if(xl <= -2) r = den_fdiv_A(Itv(xl, -2), r2);
  if(xl <= -1 && xu >= -1) r = fjoin(r, den_fdiv_C(r2));
  if(xl <= 0 && xu >= 0) r = fjoin(r, den_fdiv_B(r2));
  if(xu >= 1) r = fjoin(r, den_fdiv_A(Itv(1, xu), r2));
  return r;
}

// ---- den_cdiv (pir.hpp:577-630), flattened ----------------------------------------------------------------------
__device__ __forceinline__ Itv den_cdiv_pos(const Itv& r1, int ylo, int yhi) {
  const int xl1 = wsub(xl, 1);
  return Itv(min(cdiv(ylo, xu), cdiv(yhi, xu)), wsub(max(cdiv(ylo, xl1), cdiv(yhi, xl1)), 1));
}
__device__ __forceinline__ Itv den_cdiv_neg(const Itv& r1, int ylo, int yhi) {
  const int xl1 = wsub(xl, 1);
  return Itv(wadd(min(fdiv(ylo, xl1), fdiv(yhi, xl1)), 1), max(fdiv(ylo, xu), fdiv(yhi, xu)));
}
__device__ __forceinline__ Itv den_cdiv_A(const Itv& r1, const Itv& r2) {
  if(yl > 0) return den_cdiv_pos(r1, yl, yu);
  if(yu < 0) return den_cdiv_neg(r1, yl, yu);
  if(yl == 0 && yu == 0) return itv_bot();
  Itv r = itv_bot();
  if(yl < 0) r = den_cdiv_neg(r1, yl, -1);
  if(yu > 0) r = fjoin(r, den_cdiv_pos(r1, 1, yu));
  return r;
}
__device__ __forceinline__ Itv den_cdiv_B(const Itv& r2) {   // x = [0,0]: pir.hpp:605-608
  if(yl > 0) return Itv(LPC_MINF, wsub(wneg(yl), 1));
  if(yu < 0) return Itv(wadd(wneg(yu), 1), LPC_INF);
  return itv_top();
}
__device__ __forceinline__ Itv den_cdiv_C(const Itv& r2) {   // x = [1,1]: pir.hpp:609-615
  if(yl > 0) return Itv(yl, LPC_INF);
  if(yu < 0) return Itv(LPC_MINF, yu);
  if(0 == yl && yl < yu) return Itv(1, LPC_INF);
  if(yl < yu && yu == 0) return Itv(LPC_MINF, -1);
  if(yl == 0 && yu == 0) return itv_bot();
  return itv_top();
}
static __device__ __noinline__ Itv den_cdiv(const Itv& r1, const Itv& r2) {
  if(xl > 1 || xu < 0) return den_cdiv_A(r1, r2);
  if(xl > xu) return itv_top();
  Itv r = itv_bot();
  if(xl <= -1) r = den_cdiv_A(Itv(xl, -1), r2);
  if(xl <= 0 && xu >= 0) r = fjoin(r, den_cdiv_B(r2));
  if(xl <= 1 && xu >= 1) r = fjoin(r, den_cdiv_C(r2));
  if(xu >= 2) r = fjoin(r, den_cdiv_A(Itv(2, xu), r2));
  return r;
}

// pir.hpp:633-649
__device__ __forceinline__ Itv den_tdiv0(const Itv& r2, const Itv& r3) {   // x = [0,0]
  if(yl > 0 && zl > 0) return Itv(wadd(yl, 1), LPC_INF);
  if(yl > 0 && zu < 0) return Itv(LPC_MINF, wsub(wneg(yl), 1));
  if(yu < 0 && zl > 0) return Itv(wadd(wneg(yu), 1), LPC_INF);
  if(yu < 0 && zu < 0) return Itv(LPC_MINF, wsub(yu, 1));
  return itv_top();
}
__device__ __forceinline__ Itv den_tdiv(const Itv& r1, const Itv& r2, const Itv& r3) {
  if(xl > 0) return den_fdiv(r1, r2);
  else if(xu < 0) return den_cdiv(r1, r2);
  else if(xl == 0 && xu == 0) return den_tdiv0(r2, r3);
  else if(xl <= 0 && 0 <= xu) {
    Itv r = den_tdiv0(r2, r3);
    if(xl != 0) r = fjoin(r, den_cdiv(Itv(xl, -1), r2));
    if(xu != 0) r = fjoin(r, den_fdiv(Itv(1, xu), r2));
    return r;
  }
  return itv_top();
}
// pir.hpp:651-658
__device__ __forceinline__ Itv den_ediv(const Itv& r1, const Itv& r2, const Itv& r3) {
  if(zl > 0) return den_fdiv(r1, r2);
  else if(zu < 0) return den_cdiv(r1, r2);
  else if(zl < 0 && 0 < zu) return fjoin(den_fdiv(r1, r2), den_cdiv(r1, r2));
  return itv_top();
}

// pir.hpp:780-792 (+ itv_div_num / itv_div_den, :660-699). Out of line: divisions are rare, keep the hot
// operators' code small.
static __device__ __noinline__ void deduce_div(int op, Itv& r1, Itv& r2, Itv& r3) {
  itv_div(op, r1, r2, r3);
  if(!r1.is_bot() && !r3.is_bot()) {
    Itv n = op == D_FDIV ? num_fdiv(r1, r3) : op == D_CDIV ? num_cdiv(r1, r3)
          : op == D_TDIV ? num_tdiv(r1, r3) : num_ediv(r1, r3);
    r2.meet(n);
    if(!r2.is_bot()) {
      Itv d = op == D_FDIV ? den_fdiv(r1, r2) : op == D_CDIV ? den_cdiv(r1, r2)
            : op == D_TDIV ? den_tdiv(r1, r2, r3) : den_ediv(r1, r2, r3);
      r3.meet(d);
    }
  }
}

// The propagator x = y op z on registers: r1, r2, r3 come in as the loaded domains and leave as the domains to
// join into the store (pir.hpp:729-812). For EQ / LEQ the reference embeds explicit intervals (:730-757); they are
// folded into r1..r3 by meet, which is the same join.
template <bool HAS_DIV>
__device__ __forceinline__ void deduce_regs(int op, Itv& r1, Itv& r2, Itv& r3) {
  switch(op) {
    case D_ADD: {
      r1.lb = (yl == LPC_MINF || zl == LPC_MINF) ? xl : max(xl, wadd(yl, zl));
      r1.ub = (yu == LPC_INF || zu == LPC_INF) ? xu : min(xu, wadd(yu, zu));
      r2.lb = (xl == LPC_MINF || zu == LPC_INF) ? yl : max(yl, wsub(xl, zu));
      r2.ub = (xu == LPC_INF || zl == LPC_MINF) ? yu : min(yu, wsub(xu, zl));
      r3.lb = (xl == LPC_MINF || yu == LPC_INF) ? zl : max(zl, wsub(xl, yu));
      r3.ub = (xu == LPC_INF || yl == LPC_MINF) ? zu : min(zu, wsub(xu, yl));
      break;
    }
    case D_MUL: {
      if(yl != LPC_MINF && yu != LPC_INF && zl != LPC_MINF && zu != LPC_INF) {
        int t1 = wmul(yl, zl), t2 = wmul(yl, zu), t3 = wmul(yu, zl), t4 = wmul(yu, zu);
        r1.lb = max(xl, min(min(t1, t2), min(t3, t4)));
        r1.ub = min(xu, max(max(t1, t2), max(t3, t4)));
      }
      mul_inv(r1, r2, r3);
      mul_inv(r1, r3, r2);
      break;
    }
    case D_MIN: {
      r1.lb = max(xl, min(yl, zl));
      r1.ub = min(xu, min(yu, zu));
      r2.lb = max(yl, xl);
      if(xu < zl) r2.ub = min(yu, xu);
      r3.lb = max(zl, xl);
      if(xu < yl) r3.ub = min(zu, xu);
      break;
    }
    case D_MAX: {
      r1.lb = max(xl, max(yl, zl));
      r1.ub = min(xu, max(yu, zu));
      r2.ub = min(yu, xu);
      if(xl > zu) r2.lb = max(yl, xl);
      r3.ub = min(zu, xu);
      if(xl > yu) r3.lb = max(zl, xl);
      break;
    }
    case D_EQ: {
      if(xl == 1 && xu == 1) {
        Itv y0 = r2;
        r2.meet(r3);
        r3.meet(y0);
      }
      else if(xl == 0 && xu == 0 && (yl == yu || zl == zu)) {
        Itv s(yl == zl ? wadd(yl, 1) : LPC_MINF, yu == zu ? wsub(yu, 1) : LPC_INF);
        if(zl == zu) r2.meet(s); else r3.meet(s);
      }
      else if(yu == zl && yl == zu) r1.meet(Itv(1, 1));
      else if(yl > zu || yu < zl) r1.meet(Itv(0, 0));
      break;
    }
    case D_LEQ: {
      if(xl == 1 && xu == 1) {
        Itv s(yl, zu);
        r2.meet(s);
        r3.meet(s);
      }
      else if(xl == 0 && xu == 0) {
        Itv sy(wadd(zl, 1), yu), sz(zl, wsub(yu, 1));
        r2.meet(sy);
        r3.meet(sz);
      }
      else if(yu <= zl) r1.meet(Itv(1, 1));
      else if(yl > zu) r1.meet(Itv(0, 0));
      break;
    }
    case D_TDIV: case D_FDIV: case D_CDIV: case D_EDIV: {
      if(HAS_DIV) deduce_div(op, r1, r2, r3);
      break;
    }
    default: break;   // D_NOP padding
  }
}

// pir.hpp:417-438
__device__ __forceinline__ bool ask_regs(int op, const Itv& r1, const Itv& r2, const Itv& r3) {
  switch(op) {
    case D_EQ: return (xl == 1 && yu == zl && yl == zu) || (xu == 0 && (yu < zl || yl > zu));
    case D_LEQ: return (xl == 1 && yu <= zl) || (xu == 0 && yl > zu);
    case D_ADD: return (xl == xu && yl == yu && zl == zu && xl == wadd(yl, zl));
    case D_MUL: return xl == xu && ((yl == yu && zl == zu && xl == wmul(yl, zl))
                                 || (xl == 0 && ((yl == 0 && yu == 0) || (zl == 0 && zu == 0))));
    case D_TDIV: case D_FDIV: case D_CDIV: case D_EDIV:
      return (xl == xu && yl == yu && zl == zu && zl != 0 && xl == divop(yl, op, zl))
          || (xl == yu && xu == yl && xl == 0 && (zl > 0 || zu < 0));
    case D_MIN: return (xl == yu && xu == yl && yu <= zl) || (xl == zu && xu == zl && zu <= yl);
    case D_MAX: return (xl == yu && xu == yl && yl >= zu) || (xl == zu && xu == zl && zl >= yu);
    default: return true;   // padding never blocks extraction
  }
}

#undef xl
#undef xu
#undef yl
#undef yu
#undef zl
#undef zu

} // namespace lpc
