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

#ifdef LPC_HOST_HARNESS   // tests/native/devhost.cu runs these rules on the host
#define LPC_HD __host__ __device__ __forceinline__
#else
#define LPC_HD __device__ __forceinline__
#endif
#define LPC_INF 2147483647
#define LPC_MINF (-2147483647 - 1)

struct Itv {
  int lb, ub;
  __host__ __device__ __forceinline__ Itv() {}
  __host__ __device__ __forceinline__ Itv(int l, int u) : lb(l), ub(u) {}
  __host__ __device__ __forceinline__ bool is_bot() const { return lb > ub; }
  __host__ __device__ __forceinline__ void meet(const Itv& o) { lb = lb > o.lb ? lb : o.lb; ub = ub < o.ub ? ub : o.ub; }
};

// A FINITE bound within 2^24 of the int32 limits (lpc.h, "arithmetic"): from there on `+` and `*` may wrap.
LPC_HD bool near_inf_lo(int b) { return (unsigned)b - 0x80000001u < (1u << 24); }
LPC_HD bool near_inf_hi(int b) { return 0x7ffffffeu - (unsigned)b < (1u << 24); }

LPC_HD Itv itv_top() { return Itv(LPC_MINF, LPC_INF); }
LPC_HD Itv itv_bot() { return Itv(LPC_INF, LPC_MINF); }
// hull ignoring empty operands (lala-core Interval::join / fjoin)
LPC_HD Itv fjoin(const Itv& a, const Itv& b) {
  if(b.is_bot()) return a;
  if(a.is_bot()) return b;
  return Itv(min(a.lb, b.lb), max(a.ub, b.ub));
}

LPC_HD int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
LPC_HD int wsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
LPC_HD int wmul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }
LPC_HD int wneg(int a) { return (int)(0u - (unsigned)a); }

// One hardware division, both roundings. b == 0 -> {0,0}; b == -1 -> wrap negate (INT_MIN / -1 does not trap).
struct QR { int q, r; };
LPC_HD QR divqr(int a, int b) {
  QR o;
  if(b == 0) { o.q = 0; o.r = 0; return o; }
  if(b == -1) { o.q = wneg(a); o.r = 0; return o; }
  o.q = a / b;
  o.r = a - o.q * b;
  return o;
}
LPC_HD int fdiv_of(QR d, int b) { return (d.r != 0 && ((d.r < 0) != (b < 0))) ? d.q - 1 : d.q; }
LPC_HD int cdiv_of(QR d, int b) { return (d.r != 0 && ((d.r < 0) == (b < 0))) ? d.q + 1 : d.q; }
LPC_HD int ediv_of(QR d, int b) { return d.r >= 0 ? d.q : (b > 0 ? d.q - 1 : d.q + 1); }
LPC_HD int fdiv(int a, int b) { return fdiv_of(divqr(a, b), b); }
LPC_HD int cdiv(int a, int b) { return cdiv_of(divqr(a, b), b); }
// pir.hpp:407-415
LPC_HD int divop(int a, int op, int b) {
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
LPC_HD void mul_inv(const Itv& r1, Itv& r2, Itv& r3) {
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
    // Four corner pairs (pir.hpp:711-717); the two quotients of one divisor share its reciprocal. (Dividing only at the
    // two corners that carry the extremes - z is sign-definite here - was measured SLOWER: batch 15.5 -> 16.8 ms, config 2
    // 0.592 -> 0.623 ms on the same box; the selects cost more than the two extra quotient steps.)
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

// The division propagators live in pir_div.cuh / pir_div.cu (separate translation unit, see there).
#ifdef LPC_HOST_HARNESS
LPC_HD void deduce_div(int op, Itv& r1, Itv& r2, Itv& r3);
#else
__device__ void deduce_div(int op, Itv& r1, Itv& r2, Itv& r3);
#endif

template <bool HAS_DIV>
LPC_HD void deduce_regs(int op, Itv& r1, Itv& r2, Itv& r3) {
  switch(op) {
    case D_ADD: {
      // fast path: no infinite bound anywhere -> six fused add+min/max, no guards (identical results: the guards
      // of pir.hpp:759-764 only fire on exact INT_MIN / INT_MAX bounds)
      const int lo = min(min(xl, yl), zl), hi = max(max(xu, yu), zu);
      if(lo != LPC_MINF && hi != LPC_INF) {
        r1.lb = max(xl, wadd(yl, zl));
        r1.ub = min(xu, wadd(yu, zu));
        r2.lb = max(yl, wsub(xl, zu));
        r2.ub = min(yu, wsub(xu, zl));
        r3.lb = max(zl, wsub(xl, yu));
        r3.ub = min(zu, wsub(xu, yl));
      }
      else {
        r1.lb = (yl == LPC_MINF || zl == LPC_MINF) ? xl : max(xl, wadd(yl, zl));
        r1.ub = (yu == LPC_INF || zu == LPC_INF) ? xu : min(xu, wadd(yu, zu));
        r2.lb = (xl == LPC_MINF || zu == LPC_INF) ? yl : max(yl, wsub(xl, zu));
        r2.ub = (xu == LPC_INF || zl == LPC_MINF) ? yu : min(yu, wsub(xu, zl));
        r3.lb = (xl == LPC_MINF || yu == LPC_INF) ? zl : max(zl, wsub(xl, yu));
        r3.ub = (xu == LPC_INF || yl == LPC_MINF) ? zu : min(zu, wsub(xu, yl));
      }
      break;
    }
    case D_MUL: {
      // Both factors fixed (and small enough for an exact int32 product): x <- x meet {y*z}, and neither mul_inv can
      // move a factor - with x = {p} the corner quotients of pir.hpp:711-717 are p/z = y and p/y = z exactly, with
      // p = 0 they give 0 for the factor that must be 0; an empty x fails the store either way. Multiplication pins
      // its factors early (config 2: 99.9 % of the MUL records after one sweep), and a warp whose 32 records all take
      // this branch skips the eight divisions of the general rule.
      if(yl == yu && zl == zu && (unsigned)(yl + 46340) <= 92680u && (unsigned)(zl + 46340) <= 92680u) {
        const int p = yl * zl;
        r1.lb = max(xl, p);
        r1.ub = min(xu, p);
        break;
      }
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
LPC_HD bool ask_regs(int op, const Itv& r1, const Itv& r2, const Itv& r3) {
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
