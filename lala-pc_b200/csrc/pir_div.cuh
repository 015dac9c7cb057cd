// pir_div.cuh — the four integer-division propagators (x = y tdiv|fdiv|cdiv|ediv z) on registers.
//
// Device counterpart of pir.hpp:449-699 and :780-792. Kept apart from pir_device.cuh because it is compiled in its
// own translation unit (pir_div.cu) with LPC_DIV_FIX=3, which keeps every helper below out of line: with the helpers
// inlined into one function, CUDA 12.9 miscompiles the TDIV rules for sm_100a at -O1..-O3 (found by the exhaustive TDIV
// parity test; minimal repro and the variants tried in tools/repro/div_miscompile.cu - `-Xcicc -O0 -Xptxas -O3` is
// correct, so the fault is in the NVVM optimiser, not in ptxas as first thought; round 1 worked around it with
// `-Xptxas -O0` for the whole unit). Divisions are rare in real models and the calls cost little; the hot operators stay
// inlined and are guarded by the exhaustive GPU tests.
#pragma once
#include "pir_device.cuh"

namespace lpc {

// Negation as the division rules use it. With LPC_DIV_OPAQUE_NEG the device takes it through a call that is not inlined,
// so that ptxas cannot fold the negation into a neighbouring three-input min / max (tools/repro/div_miscompile.cu).
#if defined(LPC_DIV_OPAQUE_NEG)
static __device__ __noinline__ int dneg_call(int a) { return (int)(0u - (unsigned)a); }
LPC_HD int dneg(int a) {
#ifdef __CUDA_ARCH__
  return dneg_call(a);
#else
  return (int)(0u - (unsigned)a);
#endif
}
#else
LPC_HD int dneg(int a) { return wneg(a); }
#endif

// LPC_DIV_FIX (tools/repro/div_miscompile.cu): 0 = the code as it was, 1 = num_tdiv / den_tdiv out of line on the device,
// 2 = 1 + the zero-straddling numerator written with one maximum.
#ifndef LPC_DIV_FIX
#define LPC_DIV_FIX 0
#endif
#if LPC_DIV_FIX >= 3
#undef LPC_HD
#ifdef LPC_HOST_HARNESS
#define LPC_HD __host__ __device__ __noinline__
#else
#define LPC_HD __device__ __noinline__
#endif
#endif
#if LPC_DIV_FIX >= 1 && !defined(LPC_HOST_HARNESS)
#define LPC_DIV_TDIV_ATTR __device__ __noinline__
#elif LPC_DIV_FIX >= 1
#define LPC_DIV_TDIV_ATTR __host__ __device__ __noinline__
#else
#define LPC_DIV_TDIV_ATTR LPC_HD
#endif

#define xl r1.lb
#define xu r1.ub
#define yl r2.lb
#define yu r2.ub
#define zl r3.lb
#define zu r3.ub

// pir.hpp:449-467 — r1 = r2 / r3
LPC_HD void itv_div(int op, Itv& r1, Itv& r2, Itv& r3) {
  if(zl < 0 && zu > 0) {
    r1.lb = max(xl, min(yl, yu == LPC_MINF ? LPC_INF : dneg(yu)));
    r1.ub = min(xu, max(yl == LPC_INF ? LPC_MINF : dneg(yl), yu));
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
LPC_HD Itv num_fdiv(const Itv& r1, const Itv& r3) {
  const int xu1 = wadd(xu, 1);
  if(zl < 0 && zu > 0) {
    return Itv(min(min(xl, dneg(xu)), min(wmul(xl, zu), wadd(wmul(xu1, zl), 1))),
               max(max(dneg(xl), xu), max(wmul(xl, zl), wsub(wmul(xu1, zu), 1))));
  }
  else if(zl > 0 || zu < 0) {
    return Itv(min(min(wmul(xl, zl), wmul(xl, zu)), min(wadd(wmul(xu1, zl), 1), wadd(wmul(xu1, zu), 1))),
               max(max(wmul(xl, zl), wmul(xl, zu)), max(wsub(wmul(xu1, zl), 1), wsub(wmul(xu1, zu), 1))));
  }
  return itv_top();
}

// pir.hpp:481-491
LPC_HD Itv num_cdiv(const Itv& r1, const Itv& r3) {
  const int xl1 = wsub(xl, 1);
  if(zl < 0 && zu > 0) {
    return Itv(min(min(xl, dneg(xu)), min(wmul(xu, zl), wadd(wmul(xl1, zu), 1))),
               max(max(dneg(xl), xu), max(wmul(xu, zu), wsub(wmul(xl1, zl), 1))));
  }
  else if(zl > 0 || zu < 0) {
    return Itv(min(min(wmul(xu, zl), wmul(xu, zu)), min(wadd(wmul(xl1, zl), 1), wadd(wmul(xl1, zu), 1))),
               max(max(wmul(xu, zl), wmul(xu, zu)), max(wsub(wmul(xl1, zl), 1), wsub(wmul(xl1, zu), 1))));
  }
  return itv_top();
}

// pir.hpp:493-507
LPC_DIV_TDIV_ATTR Itv num_tdiv(const Itv& r1, const Itv& r3) {
  if(xl > 0) return num_fdiv(r1, r3);
  else if(xu < 0) return num_cdiv(r1, r3);
  else if(xl <= 0 && 0 <= xu) {
#if LPC_DIV_FIX >= 2
    // the same interval as the reference's (min(zl, -zu) + 1, max(-zl, zu) - 1), written with ONE maximum:
    // min(zl, -zu) = -max(-zl, zu)
    const int m = max(dneg(zl), zu);
    Itv r(wadd(dneg(m), 1), wsub(m, 1));
#else
    Itv r(wadd(min(zl, dneg(zu)), 1), wsub(max(dneg(zl), zu), 1));
#endif
    if(xl != 0) r = fjoin(r, num_cdiv(Itv(xl, -1), r3));
    if(xu != 0) r = fjoin(r, num_fdiv(Itv(1, xu), r3));
    return r;
  }
  return itv_top();
}

// pir.hpp:510-517
LPC_HD Itv num_ediv(const Itv& r1, const Itv& r3) {
  if(zl > 0) return num_fdiv(r1, r3);
  else if(zu < 0) return num_cdiv(r1, r3);
  else if(zl < 0 && zu > 0) return fjoin(num_cdiv(r1, Itv(zl, -1)), num_fdiv(r1, Itv(1, zu)));
  return itv_top();
}

// ---- den_fdiv (pir.hpp:520-574), flattened ----------------------------------------------------------------------
// Branch `xl > 0 || xu + 1 < 0` for y > 0 / y < 0 (pir.hpp:523-534).
LPC_HD Itv den_fdiv_pos(const Itv& r1, int ylo, int yhi) {   // y in [ylo,yhi], ylo > 0
  const int xu1 = wadd(xu, 1);
  return Itv(wadd(min(fdiv(ylo, xu1), fdiv(yhi, xu1)), 1), max(fdiv(ylo, xl), fdiv(yhi, xl)));
}
LPC_HD Itv den_fdiv_neg(const Itv& r1, int ylo, int yhi) {   // yhi < 0
  const int xu1 = wadd(xu, 1);
  return Itv(min(cdiv(ylo, xl), cdiv(yhi, xl)), wsub(max(cdiv(ylo, xu1), cdiv(yhi, xu1)), 1));
}
// x definitely outside {-1, 0}: pir.hpp:522-547
LPC_HD Itv den_fdiv_A(const Itv& r1, const Itv& r2) {
  if(yl > 0) return den_fdiv_pos(r1, yl, yu);
  if(yu < 0) return den_fdiv_neg(r1, yl, yu);
  if(yl == 0 && yu == 0) return itv_bot();
  Itv r = itv_bot();
  if(yl < 0) r = den_fdiv_neg(r1, yl, -1);
  if(yu > 0) r = fjoin(r, den_fdiv_pos(r1, 1, yu));
  return r;
}
// x = [0,0]: pir.hpp:548-552
LPC_HD Itv den_fdiv_B(const Itv& r2) {
  if(yl > 0) return Itv(wadd(yl, 1), LPC_INF);
  if(yu < 0) return Itv(LPC_MINF, wsub(yu, 1));
  return itv_top();
}
// x = [-1,-1]: pir.hpp:553-559
LPC_HD Itv den_fdiv_C(const Itv& r2) {
  if(yl > 0) return Itv(LPC_MINF, dneg(yl));
  if(yu < 0) return Itv(dneg(yu), LPC_INF);
  if(0 == yl && yl < yu) return Itv(LPC_MINF, -1);
  if(yl < yu && yu == 0) return Itv(1, LPC_INF);
  if(yl == 0 && yu == 0) return itv_bot();
  return itv_top();
}
LPC_HD Itv den_fdiv(const Itv& r1, const Itv& r2) {
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
LPC_HD Itv den_cdiv_pos(const Itv& r1, int ylo, int yhi) {
  const int xl1 = wsub(xl, 1);
  return Itv(min(cdiv(ylo, xu), cdiv(yhi, xu)), wsub(max(cdiv(ylo, xl1), cdiv(yhi, xl1)), 1));
}
LPC_HD Itv den_cdiv_neg(const Itv& r1, int ylo, int yhi) {
  const int xl1 = wsub(xl, 1);
  return Itv(wadd(min(fdiv(ylo, xl1), fdiv(yhi, xl1)), 1), max(fdiv(ylo, xu), fdiv(yhi, xu)));
}
LPC_HD Itv den_cdiv_A(const Itv& r1, const Itv& r2) {
  if(yl > 0) return den_cdiv_pos(r1, yl, yu);
  if(yu < 0) return den_cdiv_neg(r1, yl, yu);
  if(yl == 0 && yu == 0) return itv_bot();
  Itv r = itv_bot();
  if(yl < 0) r = den_cdiv_neg(r1, yl, -1);
  if(yu > 0) r = fjoin(r, den_cdiv_pos(r1, 1, yu));
  return r;
}
LPC_HD Itv den_cdiv_B(const Itv& r2) {   // x = [0,0]: pir.hpp:605-608
  if(yl > 0) return Itv(LPC_MINF, wsub(dneg(yl), 1));
  if(yu < 0) return Itv(wadd(dneg(yu), 1), LPC_INF);
  return itv_top();
}
LPC_HD Itv den_cdiv_C(const Itv& r2) {   // x = [1,1]: pir.hpp:609-615
  if(yl > 0) return Itv(yl, LPC_INF);
  if(yu < 0) return Itv(LPC_MINF, yu);
  if(0 == yl && yl < yu) return Itv(1, LPC_INF);
  if(yl < yu && yu == 0) return Itv(LPC_MINF, -1);
  if(yl == 0 && yu == 0) return itv_bot();
  return itv_top();
}
LPC_HD Itv den_cdiv(const Itv& r1, const Itv& r2) {
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
LPC_HD Itv den_tdiv0(const Itv& r2, const Itv& r3) {   // x = [0,0]
  if(yl > 0 && zl > 0) return Itv(wadd(yl, 1), LPC_INF);
  if(yl > 0 && zu < 0) return Itv(LPC_MINF, wsub(dneg(yl), 1));
  if(yu < 0 && zl > 0) return Itv(wadd(dneg(yu), 1), LPC_INF);
  if(yu < 0 && zu < 0) return Itv(LPC_MINF, wsub(yu, 1));
  return itv_top();
}
LPC_DIV_TDIV_ATTR Itv den_tdiv(const Itv& r1, const Itv& r2, const Itv& r3) {
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
LPC_HD Itv den_ediv(const Itv& r1, const Itv& r2, const Itv& r3) {
  if(zl > 0) return den_fdiv(r1, r2);
  else if(zu < 0) return den_cdiv(r1, r2);
  else if(zl < 0 && 0 < zu) return fjoin(den_fdiv(r1, r2), den_cdiv(r1, r2));
  return itv_top();
}

// pir.hpp:780-792 (+ itv_div_num / itv_div_den, :660-699).
LPC_HD void deduce_div_rules(int op, Itv& r1, Itv& r2, Itv& r3) {
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
#undef xl
#undef xu
#undef yl
#undef yu
#undef zl
#undef zu

} // namespace lpc
