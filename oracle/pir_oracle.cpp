// pir_oracle.cpp — TEST INFRASTRUCTURE ONLY. CPU restatement of the PIR hot path of lala-pc.
//
// This file is the *checker*: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load it. Nothing under lala-pc_b200/ links, imports or calls it.
//
// What it restates (paths relative to the lala-pc repository):
//   deduce(bytecode)            include/lala/pir.hpp:721-817
//   ask(bytecode)               include/lala/pir.hpp:417-438
//   div dispatch                include/lala/pir.hpp:407-415
//   itv_div, num_*, den_*       include/lala/pir.hpp:449-699
//   mul_inv                     include/lala/pir.hpp:702-718
//   EQ/LEQ [0,1] clamp          include/lala/pir.hpp:333-335
//   is_extractable's ask loop   include/lala/pir.hpp:873-884
// and, from the un-vendored dependencies (lala-core v1.2.8 / cuda-battery, CMakeLists.txt:33-37; recalled from
// the public upstream and pinned by the reference's golden vectors, see tests/test_oracle_pir.py):
//   VStore::embed  = cell.meet(u); sets the sticky bot flag when the cell becomes empty; returns changed
//   Interval       = (lb, ub) int32; top [INT_MIN, INT_MAX]; bot [INT_MAX, INT_MIN]; is_bot <=> lb > ub;
//                    meet = (max lb, min ub); join / fjoin = hull ignoring empty operands
//   battery::tdiv/fdiv/cdiv/ediv   truncated / floor / ceil / Euclidean integer division
//   GaussSeidelIteration::fixpoint(n, f[, has_changed])  do { changed = OR_i f(i) in index order } while(changed)
//
// Parity status: PINNED for PIR by the single/two-record known-answer vectors of tests/pir_test.cpp and by the
// exhaustive bounds-consistency property of tests/bound_consistency_test.hpp (see tests/test_oracle_pir.py).
//
// Arithmetic is int32 two's complement with explicit wrap-around (the reference's `yl + zl` etc. are UB on
// overflow); division is made total (b == 0 -> 0, INT_MIN / -1 -> INT_MIN) where the reference would trap. The two
// guards `xu + 1 < 0` (pir.hpp:522) and `xl - 1 > 0` (pir.hpp:579) are read as `xu < -1` / `xl > 1`, which is what
// an optimising compiler makes of them under the no-overflow assumption.

#include <cstdint>
#include <cstring>
#include <climits>
#include <vector>
#include <thread>
#include <algorithm>
#include <chrono>

namespace {

typedef int32_t v_t;
const v_t INF = INT32_MAX;
const v_t MINF = INT32_MIN;

// lala-core Sig values of the operators PIR accepts (pir.hpp:270-273); same numbering as include/lpc.h.
enum { ADD = 2, MUL = 4, MIN = 6, MAX = 7, TDIV = 25, FDIV = 27, CDIV = 29, EDIV = 31, EQ = 46, LEQ = 48 };

inline v_t wadd(v_t a, v_t b) { return (v_t)((uint32_t)a + (uint32_t)b); }
inline v_t wsub(v_t a, v_t b) { return (v_t)((uint32_t)a - (uint32_t)b); }
inline v_t wmul(v_t a, v_t b) { return (v_t)((uint32_t)a * (uint32_t)b); }
inline v_t wneg(v_t a) { return (v_t)(0u - (uint32_t)a); }
inline v_t vmin(v_t a, v_t b) { return a < b ? a : b; }
inline v_t vmax(v_t a, v_t b) { return a > b ? a : b; }

// battery::{t,f,c,e}div, made total.
inline v_t tdiv(v_t a, v_t b) {
  if(b == 0) return 0;
  if(b == -1) return wneg(a);
  return a / b;
}
inline v_t fdiv(v_t a, v_t b) {
  if(b == 0) return 0;
  if(b == -1) return wneg(a);
  v_t q = a / b, r = a % b;
  return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}
inline v_t cdiv(v_t a, v_t b) {
  if(b == 0) return 0;
  if(b == -1) return wneg(a);
  v_t q = a / b, r = a % b;
  return (r != 0 && ((r < 0) == (b < 0))) ? q + 1 : q;
}
inline v_t ediv(v_t a, v_t b) {
  if(b == 0) return 0;
  if(b == -1) return wneg(a);
  v_t q = a / b, r = a % b;
  return r >= 0 ? q : (b > 0 ? q - 1 : q + 1);
}

struct Itv {
  v_t lb, ub;
  Itv() : lb(MINF), ub(INF) {}
  Itv(v_t l, v_t u) : lb(l), ub(u) {}
  static Itv top() { return Itv(MINF, INF); }
  static Itv bot() { return Itv(INF, MINF); }
  bool is_bot() const { return lb > ub; }
  bool meet(const Itv& o) {
    bool c = false;
    if(o.lb > lb) { lb = o.lb; c = true; }
    if(o.ub < ub) { ub = o.ub; c = true; }
    return c;
  }
  // hull, empty operands ignored (lala-core Interval::join).
  void join(const Itv& o) {
    if(o.is_bot()) return;
    if(is_bot()) { *this = o; return; }
    lb = vmin(lb, o.lb);
    ub = vmax(ub, o.ub);
  }
  bool eq(v_t l, v_t u) const { return lb == l && ub == u; }
};

inline Itv fjoin(const Itv& a, const Itv& b) { Itv r(a); r.join(b); return r; }

// VStore<Interval<ZLB>>: interleaved {lb, ub} pairs + sticky bot flag.
struct Store {
  v_t* d;
  int n;
  bool bot;
  Itv get(int v) const { return Itv(d[2 * v], d[2 * v + 1]); }
  bool embed(int v, const Itv& u) {
    Itv c = get(v);
    if(c.meet(u)) {
      d[2 * v] = c.lb; d[2 * v + 1] = c.ub;
      if(c.is_bot()) bot = true;
      return true;
    }
    return false;
  }
};

struct Rec { int32_t op, x, y, z; };

inline v_t divop(v_t a, int op, v_t b) {  // pir.hpp:407-415
  switch(op) {
    case TDIV: return tdiv(a, b);
    case CDIV: return cdiv(a, b);
    case FDIV: return fdiv(a, b);
    default: return ediv(a, b);
  }
}

#define xl r1.lb
#define xu r1.ub
#define yl r2.lb
#define yu r2.ub
#define zl r3.lb
#define zu r3.ub

// pir.hpp:449-467 — r1 = r2 / r3
void itv_div(int op, Itv& r1, Itv& r2, Itv& r3) {
  if(zl < 0 && zu > 0) {
    r1.lb = vmax(xl, vmin(yl, yu == MINF ? INF : wneg(yu)));
    r1.ub = vmin(xu, vmax(yl == INF ? MINF : wneg(yl), yu));
  }
  else {
    if(zl == 0) r3.lb = 1;
    if(zu == 0) r3.ub = -1;
    if(yl == MINF || yu == INF || zl == MINF || zu == INF) return;
    if(r3.is_bot()) return;
    v_t t1 = divop(yl, op, zl), t2 = divop(yl, op, zu), t3 = divop(yu, op, zl), t4 = divop(yu, op, zu);
    r1.lb = vmax(xl, vmin(vmin(t1, t2), vmin(t3, t4)));
    r1.ub = vmin(xu, vmax(vmax(t1, t2), vmax(t3, t4)));
  }
}

// pir.hpp:469-479
Itv num_fdiv(const Itv& r1, const Itv& r3) {
  if(zl < 0 && zu > 0) {
    return Itv(vmin(vmin(xl, wneg(xu)), vmin(wmul(xl, zu), wadd(wmul(wadd(xu, 1), zl), 1))),
               vmax(vmax(wneg(xl), xu), vmax(wmul(xl, zl), wsub(wmul(wadd(xu, 1), zu), 1))));
  }
  else if(zl > 0 || zu < 0) {
    return Itv(vmin(vmin(wmul(xl, zl), wmul(xl, zu)), vmin(wadd(wmul(wadd(xu, 1), zl), 1), wadd(wmul(wadd(xu, 1), zu), 1))),
               vmax(vmax(wmul(xl, zl), wmul(xl, zu)), vmax(wsub(wmul(wadd(xu, 1), zl), 1), wsub(wmul(wadd(xu, 1), zu), 1))));
  }
  return Itv::top();
}

// pir.hpp:481-491
Itv num_cdiv(const Itv& r1, const Itv& r3) {
  if(zl < 0 && zu > 0) {
    return Itv(vmin(vmin(xl, wneg(xu)), vmin(wmul(xu, zl), wadd(wmul(wsub(xl, 1), zu), 1))),
               vmax(vmax(wneg(xl), xu), vmax(wmul(xu, zu), wsub(wmul(wsub(xl, 1), zl), 1))));
  }
  else if(zl > 0 || zu < 0) {
    return Itv(vmin(vmin(wmul(xu, zl), wmul(xu, zu)), vmin(wadd(wmul(wsub(xl, 1), zl), 1), wadd(wmul(wsub(xl, 1), zu), 1))),
               vmax(vmax(wmul(xu, zl), wmul(xu, zu)), vmax(wsub(wmul(wsub(xl, 1), zl), 1), wsub(wmul(wsub(xl, 1), zu), 1))));
  }
  return Itv::top();
}

// pir.hpp:493-507
Itv num_tdiv(const Itv& r1, const Itv& r3) {
  if(xl > 0) return num_fdiv(r1, r3);
  else if(xu < 0) return num_cdiv(r1, r3);
  else if(xl <= 0 && 0 <= xu) {
    Itv r(wadd(vmin(zl, wneg(zu)), 1), wsub(vmax(wneg(zl), zu), 1));
    if(xl != 0) r.join(num_cdiv(Itv(xl, -1), r3));
    if(xu != 0) r.join(num_fdiv(Itv(1, xu), r3));
    return r;
  }
  return Itv::top();
}

// pir.hpp:510-517
Itv num_ediv(const Itv& r1, const Itv& r3) {
  if(zl > 0) return num_fdiv(r1, r3);
  else if(zu < 0) return num_cdiv(r1, r3);
  else if(zl < 0 && zu > 0) return fjoin(num_cdiv(r1, Itv(zl, -1)), num_fdiv(r1, Itv(1, zu)));
  return Itv::top();
}

// pir.hpp:520-574
Itv den_fdiv(const Itv& r1, const Itv& r2) {
  if(xl > 0 || xu < -1) {   // `xu + 1 < 0` read without overflow
    if(yl > 0) {
      return Itv(wadd(vmin(fdiv(yl, wadd(xu, 1)), fdiv(yu, wadd(xu, 1))), 1),
                 vmax(fdiv(yl, xl), fdiv(yu, xl)));
    }
    else if(yu < 0) {
      return Itv(vmin(cdiv(yl, xl), cdiv(yu, xl)),
                 wsub(vmax(cdiv(yl, wadd(xu, 1)), cdiv(yu, wadd(xu, 1))), 1));
    }
    else if(0 == yl && yl < yu) return den_fdiv(r1, Itv(1, yu));
    else if(yl < yu && yu == 0) return den_fdiv(r1, Itv(yl, -1));
    else if(yl < 0 && 0 < yu) return fjoin(den_fdiv(r1, Itv(yl, -1)), den_fdiv(r1, Itv(1, yu)));
    else if(yl == 0 && yu == 0) return Itv::bot();
  }
  else if(xl == 0 && xu == 0) {
    if(yl > 0) return Itv(wadd(yl, 1), INF);
    else if(yu < 0) return Itv(MINF, wsub(yu, 1));
  }
  else if(xl == -1 && xu == -1) {
    if(yl > 0) return Itv(MINF, wneg(yl));
    else if(yu < 0) return Itv(wneg(yu), INF);
    else if(0 == yl && yl < yu) return Itv(MINF, -1);
    else if(yl < yu && yu == 0) return Itv(1, INF);
    else if(yl == 0 && yu == 0) return Itv::bot();
  }
  else if(xl == 0 && 0 < xu) {
    return fjoin(den_fdiv(Itv(0, 0), r2), den_fdiv(Itv(1, xu), r2));
  }
  else if(xl < -1 && xu == -1) {
    return fjoin(den_fdiv(Itv(xl, -2), r2), den_fdiv(Itv(-1, -1), r2));
  }
  else if(xl <= -1 && xu >= 0) {
    Itv r(den_fdiv(Itv(-1, -1), r2));
    r.join(den_fdiv(Itv(0, 0), r2));
    if(xl != -1) r.join(den_fdiv(Itv(xl, -2), r2));
    if(xu != 0) r.join(den_fdiv(Itv(1, xu), r2));
    return r;
  }
  return Itv::top();
}

// pir.hpp:577-630
Itv den_cdiv(const Itv& r1, const Itv& r2) {
  if(xl > 1 || xu < 0) {   // `xl - 1 > 0` read without overflow
    if(yl > 0) {
      return Itv(vmin(cdiv(yl, xu), cdiv(yu, xu)),
                 wsub(vmax(cdiv(yl, wsub(xl, 1)), cdiv(yu, wsub(xl, 1))), 1));
    }
    else if(yu < 0) {
      return Itv(wadd(vmin(fdiv(yl, wsub(xl, 1)), fdiv(yu, wsub(xl, 1))), 1),
                 vmax(fdiv(yl, xu), fdiv(yu, xu)));
    }
    else if(0 == yl && yl < yu) return den_cdiv(r1, Itv(1, yu));
    else if(yl < yu && yu == 0) return den_cdiv(r1, Itv(yl, -1));
    else if(yl < 0 && 0 < yu) return fjoin(den_cdiv(r1, Itv(yl, -1)), den_cdiv(r1, Itv(1, yu)));
    else if(yl == 0 && yu == 0) return Itv::bot();
  }
  else if(xl == 0 && xu == 0) {
    if(yl > 0) return Itv(MINF, wsub(wneg(yl), 1));
    else if(yu < 0) return Itv(wadd(wneg(yu), 1), INF);
  }
  else if(xl == 1 && xu == 1) {
    if(yl > 0) return Itv(yl, INF);
    else if(yu < 0) return Itv(MINF, yu);
    else if(0 == yl && yl < yu) return Itv(1, INF);
    else if(yl < yu && yu == 0) return Itv(MINF, -1);
    else if(yl == 0 && yu == 0) return Itv::bot();
  }
  else if(xl < 0 && xu == 0) {
    return fjoin(den_cdiv(Itv(xl, -1), r2), den_cdiv(Itv(0, 0), r2));
  }
  else if(xl == 1 && 1 < xu) {
    return fjoin(den_cdiv(Itv(1, 1), r2), den_cdiv(Itv(2, xu), r2));
  }
  else if(xl <= 0 && xu >= 1) {
    Itv r(den_cdiv(Itv(1, 1), r2));
    r.join(den_cdiv(Itv(0, 0), r2));
    if(xl != 0) r.join(den_cdiv(Itv(xl, -1), r2));
    if(xu != 1) r.join(den_cdiv(Itv(2, xu), r2));
    return r;
  }
  return Itv::top();
}

// pir.hpp:633-649
Itv den_tdiv(const Itv& r1, const Itv& r2, const Itv& r3) {
  if(xl > 0) return den_fdiv(r1, r2);
  else if(xu < 0) return den_cdiv(r1, r2);
  else if(xl == 0 && xu == 0) {
    if(yl > 0 && zl > 0) return Itv(wadd(yl, 1), INF);
    if(yl > 0 && zu < 0) return Itv(MINF, wsub(wneg(yl), 1));
    if(yu < 0 && zl > 0) return Itv(wadd(wneg(yu), 1), INF);
    if(yu < 0 && zu < 0) return Itv(MINF, wsub(yu, 1));
  }
  else if(xl <= 0 && 0 <= xu) {
    Itv r(den_tdiv(Itv(0, 0), r2, r3));
    if(xl != 0) r.join(den_cdiv(Itv(xl, -1), r2));
    if(xu != 0) r.join(den_fdiv(Itv(1, xu), r2));
    return r;
  }
  return Itv::top();
}

// pir.hpp:651-658
Itv den_ediv(const Itv& r1, const Itv& r2, const Itv& r3) {
  if(zl > 0) return den_fdiv(r1, r2);
  else if(zu < 0) return den_cdiv(r1, r2);
  else if(zl < 0 && 0 < zu) return fjoin(den_fdiv(r1, r2), den_cdiv(r1, r2));
  return Itv::top();
}

// pir.hpp:660-699
void itv_div_num(int op, Itv& r1, Itv& r2, Itv& r3) {
  switch(op) {
    case FDIV: r2.meet(num_fdiv(r1, r3)); break;
    case CDIV: r2.meet(num_cdiv(r1, r3)); break;
    case TDIV: r2.meet(num_tdiv(r1, r3)); break;
    case EDIV: r2.meet(num_ediv(r1, r3)); break;
  }
}
void itv_div_den(int op, Itv& r1, Itv& r2, Itv& r3) {
  switch(op) {
    case FDIV: r3.meet(den_fdiv(r1, r2)); break;
    case CDIV: r3.meet(den_cdiv(r1, r2)); break;
    case TDIV: r3.meet(den_tdiv(r1, r2, r3)); break;
    case EDIV: r3.meet(den_ediv(r1, r2, r3)); break;
  }
}

// pir.hpp:702-718
void mul_inv(const Itv& r1, Itv& r2, Itv& r3) {
  if(xl > 0 || xu < 0) {
    if(zl == 0) r3.lb = 1;
    if(zu == 0) r3.ub = -1;
  }
  if((xl > 0 || xu < 0) && zl < 0 && zu > 0) {
    r2.lb = vmax(yl, vmin(xl, xu == MINF ? INF : wneg(xu)));
    r2.ub = vmin(yu, vmax(xl == INF ? MINF : wneg(xl), xu));
  }
  else if(xl > 0 || xu < 0 || zl > 0 || zu < 0) {
    if(xl == MINF || xu == INF || zl == MINF || zu == INF) return;
    if(r3.is_bot()) return;
    r2.lb = vmax(yl, vmin(vmin(cdiv(xl, zl), cdiv(xl, zu)), vmin(cdiv(xu, zl), cdiv(xu, zu))));
    r2.ub = vmin(yu, vmax(vmax(fdiv(xl, zl), fdiv(xl, zu)), vmax(fdiv(xu, zl), fdiv(xu, zu))));
  }
}

// pir.hpp:721-817
bool deduce(Store& s, const Rec& b) {
  bool has_changed = false;
  Itv r1 = s.get(b.x), r2 = s.get(b.y), r3 = s.get(b.z);
  switch(b.op) {
    case EQ: {
      if(r1.eq(1, 1)) {
        has_changed |= s.embed(b.y, r3);
        has_changed |= s.embed(b.z, r2);
      }
      else if(r1.eq(0, 0) && (yl == yu || zl == zu)) {
        has_changed |= s.embed(zl == zu ? b.y : b.z,
          Itv(yl == zl ? wadd(yl, 1) : MINF, yu == zu ? wsub(yu, 1) : INF));
      }
      else if(yu == zl && yl == zu) has_changed |= s.embed(b.x, Itv(1, 1));
      else if(yl > zu || yu < zl) has_changed |= s.embed(b.x, Itv(0, 0));
      return has_changed;
    }
    case LEQ: {
      if(r1.eq(1, 1)) {
        has_changed |= s.embed(b.y, Itv(yl, zu));
        has_changed |= s.embed(b.z, Itv(yl, zu));
      }
      else if(r1.eq(0, 0)) {
        has_changed |= s.embed(b.y, Itv(wadd(zl, 1), yu));
        has_changed |= s.embed(b.z, Itv(zl, wsub(yu, 1)));
      }
      else if(yu <= zl) has_changed |= s.embed(b.x, Itv(1, 1));
      else if(yl > zu) has_changed |= s.embed(b.x, Itv(0, 0));
      return has_changed;
    }
    case ADD: {
      r1.lb = (yl == MINF || zl == MINF) ? xl : vmax(xl, wadd(yl, zl));
      r1.ub = (yu == INF || zu == INF) ? xu : vmin(xu, wadd(yu, zu));
      r2.lb = (xl == MINF || zu == INF) ? yl : vmax(yl, wsub(xl, zu));
      r2.ub = (xu == INF || zl == MINF) ? yu : vmin(yu, wsub(xu, zl));
      r3.lb = (xl == MINF || yu == INF) ? zl : vmax(zl, wsub(xl, yu));
      r3.ub = (xu == INF || yl == MINF) ? zu : vmin(zu, wsub(xu, yl));
      break;
    }
    case MUL: {
      if(yl != MINF && yu != INF && zl != MINF && zu != INF) {
        v_t t1 = wmul(yl, zl), t2 = wmul(yl, zu), t3 = wmul(yu, zl), t4 = wmul(yu, zu);
        r1.lb = vmax(xl, vmin(vmin(t1, t2), vmin(t3, t4)));
        r1.ub = vmin(xu, vmax(vmax(t1, t2), vmax(t3, t4)));
      }
      mul_inv(r1, r2, r3);
      mul_inv(r1, r3, r2);
      break;
    }
    case TDIV: case CDIV: case FDIV: case EDIV: {
      itv_div(b.op, r1, r2, r3);
      if(!r1.is_bot() && !r3.is_bot()) {
        itv_div_num(b.op, r1, r2, r3);
        if(!r2.is_bot()) itv_div_den(b.op, r1, r2, r3);
      }
      break;
    }
    case MIN: {
      r1.lb = vmax(xl, vmin(yl, zl));
      r1.ub = vmin(xu, vmin(yu, zu));
      r2.lb = vmax(yl, xl);
      if(xu < zl) r2.ub = vmin(yu, xu);
      r3.lb = vmax(zl, xl);
      if(xu < yl) r3.ub = vmin(zu, xu);
      break;
    }
    case MAX: {
      r1.lb = vmax(xl, vmax(yl, zl));
      r1.ub = vmin(xu, vmax(yu, zu));
      r2.ub = vmin(yu, xu);
      if(xl > zu) r2.lb = vmax(yl, xl);
      r3.ub = vmin(zu, xu);
      if(xl > yu) r3.lb = vmax(zl, xl);
      break;
    }
    default: return false;
  }
  has_changed |= s.embed(b.x, r1);
  has_changed |= s.embed(b.y, r2);
  has_changed |= s.embed(b.z, r3);
  return has_changed;
}

// pir.hpp:417-438
bool ask(const Store& s, const Rec& b) {
  Itv r1 = s.get(b.x), r2 = s.get(b.y), r3 = s.get(b.z);
  switch(b.op) {
    case EQ: return (xl == 1 && yu == zl && yl == zu) || (xu == 0 && (yu < zl || yl > zu));
    case LEQ: return (xl == 1 && yu <= zl) || (xu == 0 && yl > zu);
    case ADD: return (xl == xu && yl == yu && zl == zu && xl == wadd(yl, zl));
    case MUL: return xl == xu &&
                ((yl == yu && zl == zu && xl == wmul(yl, zl))
              || (xl == 0 && (r2.eq(0, 0) || r3.eq(0, 0))));
    case TDIV: case CDIV: case FDIV: case EDIV:
      return (xl == xu && yl == yu && zl == zu && zl != 0 && xl == divop(yl, b.op, zl))
          || (xl == yu && xu == yl && xl == 0 && (zl > 0 || zu < 0));
    case MIN: return (xl == yu && xu == yl && yu <= zl) || (xl == zu && xu == zl && zu <= yl);
    case MAX: return (xl == yu && xu == yl && yl >= zu) || (xl == zu && xu == zl && zl >= yu);
    default: return false;
  }
}

#undef xl
#undef xu
#undef yl
#undef yu
#undef zl
#undef zu

bool scan_bot(const v_t* d, int n) {
  for(int i = 0; i < n; ++i) if(d[2 * i] > d[2 * i + 1]) return true;
  return false;
}

struct Stats { int32_t has_changed, is_bot; int64_t sweeps, deductions; double seconds; };

// GaussSeidelIteration::fixpoint(n, f, has_changed) as called at tests/pir_test.cpp:60-62, 82-86.
// stop_on_bot = the `must_stop` overload with must_stop = is_bot (checked once per sweep).
void gauss_seidel(Store& s, const Rec* recs, int64_t n, int stop_on_bot, int64_t max_sweeps, Stats* st) {
  auto t0 = std::chrono::steady_clock::now();
  bool changed = true, any = false;
  int64_t sweeps = 0;
  while(changed && !(stop_on_bot && s.bot) && (max_sweeps <= 0 || sweeps < max_sweeps)) {
    changed = false;
    for(int64_t i = 0; i < n; ++i) changed |= deduce(s, recs[i]);
    any |= changed;
    ++sweeps;
  }
  auto t1 = std::chrono::steady_clock::now();
  if(st) {
    st->has_changed = any; st->is_bot = s.bot; st->sweeps = sweeps; st->deductions = sweeps * n;
    st->seconds = std::chrono::duration<double>(t1 - t0).count();
  }
}

} // namespace

extern "C" {

struct lpco_stats { int32_t has_changed, is_bot; int64_t sweeps, deductions; double seconds; };

// One deduce step on a store given as interleaved {lb,ub} pairs. Returns changed; *is_bot is sticky (in/out).
int lpco_pir_deduce(int32_t* lbub, int32_t nvars, const int32_t* rec4, int32_t* is_bot) {
  Store s{lbub, nvars, is_bot && *is_bot};
  Rec r{rec4[0], rec4[1], rec4[2], rec4[3]};
  bool c = deduce(s, r);
  if(is_bot) *is_bot = s.bot;
  return c;
}

int lpco_pir_ask(const int32_t* lbub, int32_t nvars, const int32_t* rec4) {
  Store s{const_cast<int32_t*>(lbub), nvars, false};
  Rec r{rec4[0], rec4[1], rec4[2], rec4[3]};
  return ask(s, r);
}

// pir.hpp:333-335
void lpco_pir_clamp_reified(int32_t* lbub, int32_t nvars, const int32_t* recs, int64_t n, int32_t* is_bot) {
  Store s{lbub, nvars, is_bot && *is_bot};
  const Rec* r = reinterpret_cast<const Rec*>(recs);
  for(int64_t i = 0; i < n; ++i)
    if(r[i].op == EQ || r[i].op == LEQ) s.embed(r[i].x, Itv(0, 1));
  if(is_bot) *is_bot = s.bot;
}

// One Gauss-Seidel fixpoint. The bot flag starts from a scan of the store (a store created with an empty
// variable is at bot before the first sweep, bound_consistency_test.hpp:22-25).
void lpco_pir_fixpoint(int32_t* lbub, int32_t nvars, const int32_t* recs, int64_t n, int32_t stop_on_bot,
                       int64_t max_sweeps, lpco_stats* out) {
  Store s{lbub, nvars, scan_bot(lbub, nvars)};
  Stats st;
  gauss_seidel(s, reinterpret_cast<const Rec*>(recs), n, stop_on_bot, max_sweeps, &st);
  if(out) { out->has_changed = st.has_changed; out->is_bot = st.is_bot; out->sweeps = st.sweeps;
            out->deductions = st.deductions; out->seconds = st.seconds; }
}

// Chaotic iteration in a caller-given order (perm of 0..n-1 applied every sweep): used to test that the
// fixpoint does not depend on the schedule.
void lpco_pir_fixpoint_perm(int32_t* lbub, int32_t nvars, const int32_t* recs, int64_t n, const int64_t* perm,
                            int32_t stop_on_bot, int64_t max_sweeps, lpco_stats* out) {
  Store s{lbub, nvars, scan_bot(lbub, nvars)};
  const Rec* r = reinterpret_cast<const Rec*>(recs);
  bool changed = true, any = false;
  int64_t sweeps = 0;
  while(changed && !(stop_on_bot && s.bot) && (max_sweeps <= 0 || sweeps < max_sweeps)) {
    changed = false;
    for(int64_t i = 0; i < n; ++i) changed |= deduce(s, r[perm[i]]);
    any |= changed;
    ++sweeps;
  }
  if(out) { out->has_changed = any; out->is_bot = s.bot; out->sweeps = sweeps; out->deductions = sweeps * n; out->seconds = 0; }
}

// The ask loop of is_extractable (pir.hpp:873-884). bits may be null.
int64_t lpco_pir_ask_all(const int32_t* lbub, int32_t nvars, const int32_t* recs, int64_t n, uint8_t* bits) {
  Store s{const_cast<int32_t*>(lbub), nvars, false};
  const Rec* r = reinterpret_cast<const Rec*>(recs);
  int64_t c = 0;
  for(int64_t i = 0; i < n; ++i) { bool e = ask(s, r[i]); if(bits) bits[i] = e; c += e; }
  return c;
}

// Batched mode on the host: n_stores independent stores [n_stores][nvars][2] over one shared table, static
// partition over `threads` std::threads. Per store: flags bit0 = bot, bit1 = all entailed; sweeps.
// Returns wall seconds.
double lpco_pir_batch_fixpoint(int32_t* lbub, int32_t n_stores, int32_t nvars, const int32_t* recs, int64_t n,
                               int32_t threads, uint8_t* flags, int32_t* sweeps, int64_t* deductions) {
  const Rec* r = reinterpret_cast<const Rec*>(recs);
  if(threads < 1) threads = 1;
  std::vector<int64_t> ded(threads, 0);
  auto t0 = std::chrono::steady_clock::now();
  auto work = [&](int tid) {
    int64_t lo = (int64_t)n_stores * tid / threads, hi = (int64_t)n_stores * (tid + 1) / threads;
    for(int64_t k = lo; k < hi; ++k) {
      int32_t* d = lbub + k * 2 * (int64_t)nvars;
      Store s{d, nvars, scan_bot(d, nvars)};
      Stats st;
      gauss_seidel(s, r, n, 1, 0, &st);
      ded[tid] += st.deductions;
      uint8_t f = s.bot ? 1 : 0;
      if(!s.bot) {
        bool all = true;
        for(int64_t i = 0; i < n && all; ++i) all = ask(s, r[i]);
        if(all) f |= 2;
      }
      if(flags) flags[k] = f;
      if(sweeps) sweeps[k] = (int32_t)st.sweeps;
    }
  };
  std::vector<std::thread> th;
  for(int t = 1; t < threads; ++t) th.emplace_back(work, t);
  work(0);
  for(auto& t : th) t.join();
  auto t1 = std::chrono::steady_clock::now();
  if(deductions) { int64_t tot = 0; for(auto d : ded) tot += d; *deductions = tot; }
  return std::chrono::duration<double>(t1 - t0).count();
}

// Depth-first search around the fixpoint with snapshot / restore (pir.hpp:857-870): the checker of lpc_batch_search.
// Variable = first non-singleton of `bvars` (input order), split = bisection with the lower half first, leaf = all
// branching variables fixed, solution = leaf on which every propagator is entailed (is_extractable, pir.hpp:873-884).
// The search strategy itself is lala-core's (un-vendored SearchTree / split strategies): PARITY UNPINNED for the order of
// exploration; the counts checked here do not depend on it as long as the tree is explored completely.
// out: n_stores x 6 int64 {solutions, nodes, fails, best, incomplete, unknown_leaves}. roots are not modified.
void lpco_pir_search(const int32_t* roots, int32_t n_stores, int32_t nvars, const int32_t* recs, int64_t n,
                     const int32_t* bvars, int32_t nb, int32_t objective_var, int64_t max_nodes, int32_t max_depth,
                     int32_t threads, int64_t* out) {
  const Rec* r = reinterpret_cast<const Rec*>(recs);
  if(threads < 1) threads = 1;
  auto work = [&](int tid) {
    std::vector<int32_t> cur(2 * (size_t)nvars);
    std::vector<std::vector<int32_t>> stack;
    for(int64_t k = tid; k < n_stores; k += threads) {
      std::copy(roots + k * 2 * (int64_t)nvars, roots + (k + 1) * 2 * (int64_t)nvars, cur.begin());
      stack.clear();
      int64_t sol = 0, nodes = 0, fails = 0, unk = 0, best = INT32_MAX, incomplete = 0;
      while(true) {
        Store s{cur.data(), nvars, scan_bot(cur.data(), nvars)};
        Stats st;
        gauss_seidel(s, r, n, 1, 0, &st);
        ++nodes;
        bool backtrack = false;
        if(s.bot) { ++fails; backtrack = true; }
        else {
          int idx = -1;
          for(int i = 0; i < nb; ++i) if(cur[2 * bvars[i]] < cur[2 * bvars[i] + 1]) { idx = i; break; }
          if(idx < 0) {
            bool all = true;
            for(int64_t i = 0; i < n && all; ++i) all = ask(s, r[i]);
            if(all) { ++sol; if(objective_var >= 0 && cur[2 * objective_var] < best) best = cur[2 * objective_var]; }
            else ++unk;
            backtrack = true;
          }
          else if((int)stack.size() >= max_depth || (max_nodes > 0 && nodes >= max_nodes)) { incomplete = 1; break; }
          else {
            const int v = bvars[idx];
            const int64_t lb = cur[2 * v], ub = cur[2 * v + 1];
            const int32_t mid = (int32_t)(lb + ((ub - lb) >> 1));
            stack.push_back(cur);                       // snapshot of the right branch
            stack.back()[2 * v] = mid + 1;
            cur[2 * v + 1] = mid;
          }
        }
        if(backtrack) {
          if(stack.empty()) break;
          cur = stack.back();                           // restore
          stack.pop_back();
        }
      }
      int64_t* o = out + k * 6;
      o[0] = sol; o[1] = nodes; o[2] = fails; o[3] = best; o[4] = incomplete; o[5] = unk;
    }
  };
  std::vector<std::thread> th;
  for(int t = 1; t < threads; ++t) th.emplace_back(work, t);
  work(0);
  for(auto& t : th) t.join();
}

// battery division helpers, exported for the division known-answer tests.
int32_t lpco_div(int32_t a, int32_t op, int32_t b) { return divop(a, op, b); }

} // extern "C"
