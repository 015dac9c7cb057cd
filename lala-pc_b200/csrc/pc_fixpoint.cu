// pc_fixpoint.cu — the PC fixpoint over flattened n-ary propagators (include/lpc_pc.h) for sm_100a.
//
// Replaces  GaussSeidelIteration{}.fixpoint(pc.num_deductions(), [&](size_t i){ return pc.deduce(i); }, has_changed)
// (tests/pc_test.cpp:91-94) where pc.deduce(i) walks a heap-allocated formula / term tree (pc.hpp:671-680). The flat
// table is streamed instead, ONE LANE PER TERM: the builder packs whole propagators into 32-lane tiles (a 16-byte
// {coef, var, meta, rhs} record per lane, a reified sum gets one more lane for its Boolean), a warp fetches a tile with
// one coalesced 512-byte load, every lane gathers its own domain from the L2-resident store, the lanes of a propagator
// combine through segmented shuffles (sum of the term bounds, partner domain, ballot of refuted literals) and each lane
// joins its own variable by atomicMax / atomicMin (atomicAnd on bitset stores). Two dependent memory round trips per
// propagator whatever its arity, where a thread walking its terms one after the other pays two per term.
// Propagators whose arithmetic could leave int32 (an infinite or empty operand, |term bound| >= 2^26) and propagators
// with more than 32 lanes are evaluated by the term-by-term routine of pc_device.cuh (`pc_deduce`), which is also what
// PC::deduce(i) uses; both compute the same monotone operator, so the fixpoint is the same.
// Same persistent cooperative skeleton as the PIR kernel: dense chaotic sweeps, block-voted has_changed / bot, a grid
// barrier per sweep.
#include "lpc_internal.cuh"
#include "pc_device.cuh"
#include "pc_tree.cuh"   // tree_check (host) for the table builder
#include "grid_barrier.cuh"

#include <algorithm>
#include <cstring>

#include "../../include/lpc_pc.h"
#include <mutex>

namespace lpc {

constexpr int PC_TPB = 256;
// kernel variants <tiles whose loads a warp keeps in flight together (0 = software pipeline), min blocks per SM>; LPC_PC_VARIANT picks one for
// tuning runs (tools/pc_probe.py), the default is the fastest measured on config 3 / config 5
constexpr int PC_NVAR = 5;
constexpr int PC_VAR_DEFAULT = 0;

template <bool BITS> struct PcAcc { typedef GlobalAcc type; };
template <> struct PcAcc<true> { typedef GlobalBitAcc type; };
template <bool BITS, class Acc> __device__ __noinline__ int pc_step(Acc& acc, const int4 h, const int2* terms) {
  if constexpr(BITS) return pc_deduce_bits(acc, h, terms);
  else return pc_deduce(acc, h, terms);
}
__device__ __forceinline__ GlobalAcc make_acc(int2* store, GlobalAcc*) { return GlobalAcc{store, 0, 0}; }
__device__ __forceinline__ GlobalBitAcc make_acc(int2* store, GlobalBitAcc*) { return GlobalBitAcc{reinterpret_cast<u64*>(store), 0, 0}; }

// ---- one tile = one warp step ------------------------------------------------------------------------------------------
// Join of one variable given the domain this lane read (VStore::embed): bit0 = tightened, bit1 = became empty.
__device__ __forceinline__ int lane_embed(int2* s, int v, const Itv& old, const Itv& u) {
  int f = 0;
  if(u.lb > old.lb) { atomicMax(&s[v].x, u.lb); f = 1; }
  if(u.ub < old.ub) { atomicMin(&s[v].y, u.ub); f = 1; }
  if(f && max(u.lb, old.lb) > min(u.ub, old.ub)) f |= 2;
  return f;
}
__device__ __forceinline__ int lane_embed_bits(u64* s, int v, u64 old, u64 u) {
  const u64 nw = old & u;
  if(nw == old) return 0;
  atomicAnd(&s[v], u);
  return nw == 0 ? 3 : 1;
}
// Euclidean quotient of int32 operands that cannot overflow (|a| < 2^31, b != 0 finite)
__device__ __forceinline__ int ediv_small(int a, int b) {
  if(b == 1) return a;
  int q = a / b;
  const int r = a - q * b;
  if(r < 0) q += b > 0 ? -1 : 1;
  return q;
}
// x with the value p removed (Equality<true>::deduce, formula.hpp:645-652)
__device__ __forceinline__ Itv itv_shave(const Itv& x, int p) {
  Itv lo = x, hi = x;
  lo.meet(Itv(b_add(p, 1), LPC_INF));
  hi.meet(Itv(LPC_MINF, b_sub(p, 1)));
  return fjoin(lo, hi);
}

// the rare path of a tile: one propagator evaluated term by term (kept out of line: it would otherwise set the register
// count of the whole kernel)
template <class Acc>
__device__ __forceinline__ int tile_fallback(const PcTableDev& t, Acc& acc, int p) {
  const int4 h = t.hdr[p];
  return pc_step<false>(acc, h, t.terms + h.y);
}

// `L` = this lane's record of the tile, `raw` = the 8-byte cell of its variable (already loaded by the caller so that
// several tiles' loads are in flight together).
template <bool BITS, class Acc>
__device__ __forceinline__ int tile_step(const PcTableDev& t, Acc& acc, int2* store, long long tile, int lane, const int4 L,
                                         const int2 raw) {
  const unsigned FULL = 0xffffffffu;
  const int coef = L.x, var = L.y, meta = L.z, rhs = L.w;
  const int kind = PC_META_KIND(meta), seg0 = meta & 31, len = (meta >> 5) & 63, pos = lane - seg0;
  const bool active = kind != 0, isb = PC_META_EXTRA(meta);
  const int last = seg0 + len - 1;
  const unsigned segmask = (len >= 32 ? FULL : ((1u << len) - 1u)) << seg0;
  // the lane whose domain this lane needs: the Boolean of a reified sum, else the other side of a binary propagator
  const int partner = pc_is_linear(kind) ? last : (len == 2 ? seg0 + 1 - pos : lane);
  int f = 0;
  if constexpr(BITS) {
    u64* cells = reinterpret_cast<u64*>(store);
    const u64 dom = active ? ((u64)(unsigned)raw.x | ((u64)(unsigned)raw.y << 32)) : ~0ull;
    const u64 pd = __shfl_sync(FULL, dom, partner);
    const bool refuted = kind == PC_CLAUSE && nb_lit_ask(coef > 0, dom);
    const unsigned open = __ballot_sync(FULL, kind == PC_CLAUSE && !refuted) & segmask;
    if(!active) return 0;
    if(dom == 0) f |= 2;
    switch(kind) {
      case PC_EQ: f |= lane_embed_bits(cells, var, dom, pd); break;
      case PC_NEQ:
        if(len == 2) { if(nb_singleton(pd)) f |= lane_embed_bits(cells, var, dom, ~pd); }
        else { const u64 r = nb_range(rhs, rhs); if(nb_singleton(r)) f |= lane_embed_bits(cells, var, dom, ~r); }
        break;
      case PC_CLAUSE: {
        const int n_open = __popc(open);
        if((n_open == 1 && !refuted) || (n_open == 0 && lane == last)) f |= lane_embed_bits(cells, var, dom, coef < 0 ? 2ull : 4ull);
        break;
      }
      case PC_ABS_EQ:
        if(pos == 0) f |= lane_embed_bits(cells, var, dom, pd | nb_neg(pd));   // x <- y join -y
        else f |= lane_embed_bits(cells, var, dom, nb_abs(pd));                // y <- |x|
        break;
      default: break;
    }
    return f;
  }
  else {
    const Itv dom = active ? Itv(raw.x, raw.y) : Itv(0, 0);
    const bool lin = pc_is_linear(kind);
    const bool term_lane = lin && !isb;
    // c * x as the hull of the two products (terms.hpp:399-405). A propagator is "tame" when every operand is non-empty
    // and every term bound is below 2^24 in magnitude: with at most 32 lanes and |rhs| < 2^30 (checked by the table
    // builder) nothing below can leave int32, so plain integer arithmetic gives what the saturating bound arithmetic
    // of pc_device.cuh gives. Infinite bounds fail the test by their size.
    // tiles without a linear propagator (config 5: =, !=, clauses, abs) skip the products, the tameness vote and the scan
    const int maxlen = __reduce_max_sync(FULL, lin ? len : 0);
    int ti_lb = 0, ti_ub = 0, all_lb = 0, all_ub = 0;
    unsigned wildm = 0, heads = 0;
    if(maxlen > 0) {
      const long long p0 = (long long)coef * dom.lb, p1 = (long long)coef * dom.ub;
      const long long tlo = min(p0, p1), thi = max(p0, p1);
      const bool small_or_inf = (dom.lb == LPC_MINF || dom.lb > -(1 << 30)) && (dom.ub == LPC_INF || dom.ub < (1 << 30));
      const bool tame = dom.lb <= dom.ub && (term_lane ? (tlo > -(1 << 24) && thi < (1 << 24)) : small_or_inf);
      ti_lb = term_lane ? (int)tlo : 0; ti_ub = term_lane ? (int)thi : 0;
      wildm = __ballot_sync(FULL, active && lin && !tame);
      heads = __ballot_sync(FULL, active && pos == 0);
      // Segmented inclusive sums of the term bounds, then the propagator's total from its last lane. Four doubling steps,
      // unrolled, cover propagators of up to 16 lanes (the loop over a run-time step count cost 113 of the 300
      // instructions of a tile step in the profile: shuffles, selects, loop control and a convergence check per shuffle);
      // a fifth step runs only for tiles that hold a longer propagator. (A masked warp reduction per propagator,
      // __reduce_add_sync, was tried: REDUX delivers ONE result per warp, so lane-varying masks fall back to a loop over
      // the distinct masks - slower, 119 vs 97 us on config 3.)
      int slb = ti_lb, sub_ = ti_ub;
#pragma unroll
      for(int off = 1; off <= 8; off <<= 1) {
        const int a = __shfl_up_sync(FULL, slb, off), b = __shfl_up_sync(FULL, sub_, off);
        const bool in = pos >= off;
        slb = wadd(slb, in ? a : 0); sub_ = wadd(sub_, in ? b : 0);
      }
      if(maxlen > 16) {
        const int a = __shfl_up_sync(FULL, slb, 16), b = __shfl_up_sync(FULL, sub_, 16);
        const bool in = pos >= 16;
        slb = wadd(slb, in ? a : 0); sub_ = wadd(sub_, in ? b : 0);
      }
      all_lb = __shfl_sync(FULL, slb, last); all_ub = __shfl_sync(FULL, sub_, last);
    }
    const Itv pd(__shfl_sync(FULL, dom.lb, partner), __shfl_sync(FULL, dom.ub, partner));
    const bool refuted = kind == PC_CLAUSE && lit_ask(coef > 0, dom);
    const unsigned open = __ballot_sync(FULL, kind == PC_CLAUSE && !refuted) & segmask;
    if(!active) return 0;
    if(dom.lb > dom.ub) f |= 2;
    if(lin) {
      if(wildm & segmask) {   // rare: the term-by-term routine, by the first lane of the propagator
        if(pos == 0) f |= tile_fallback(t, acc, t.tile_prop0[tile] + __popc(heads & ((1u << lane) - 1u)));
        return f;
      }
      // The interval u the sum is met with (Inequality / Equality / Biconditional::deduce, formula.hpp:796-805,
      // 672-681, 421-427); pd = the domain of the extra lane (Boolean / z). An infinite end of u constrains nothing.
      int ulb = LPC_MINF, uub = LPC_INF;
      bool to_terms = true;
      switch(kind) {
        case PC_LIN_LE: uub = rhs; break;
        case PC_LIN_GE: ulb = rhs; break;
        case PC_LIN_GT: ulb = rhs + 1; break;
        case PC_LIN_EQ: ulb = rhs; uub = rhs; break;
        case PC_LIN_EQ_VAR:   // z <- sum first (its own lane), then the sum is met with the new z
          ulb = max(pd.lb, all_lb); uub = min(pd.ub, all_ub);
          if(isb) f |= lane_embed(store, var, dom, Itv(all_lb, all_ub));
          break;
        default:   // PC_REIF_LIN_LE
          if(pd.lb > 0 || pd.ub < 0) uub = rhs;                  // b true (b does not contain 0)
          else if(pd.lb == 0 && pd.ub == 0) ulb = rhs + 1;       // b false
          else {
            to_terms = false;
            if(isb) {
              if(all_ub <= rhs) f |= lane_embed(store, var, dom, Itv(1, 1));
              else if(all_lb > rhs) f |= lane_embed(store, var, dom, Itv(0, 0));
            }
          }
      }
      if(to_terms && !isb) {
        // Nary<Add>::embed (terms.hpp:480-499): c * x in [ulb - (all.ub - t.ub), uub - (all.lb - t.lb)], then x's side
        // by Euclidean division, as the hull of both ends (GroupMul::left_residual, terms.hpp:249-253): a floor for
        // c > 0, a ceiling for c < 0, an infinite end staying infinite with the sign of c.
        const bool has_lo = ulb != LPC_MINF, has_hi = uub != LPC_INF;
        int p = coef > 0 ? LPC_MINF : LPC_INF, q = coef > 0 ? LPC_INF : LPC_MINF;
        if(has_lo) p = ediv_small(ulb - (all_ub - ti_ub), coef);
        if(has_hi) q = ediv_small(uub - (all_lb - ti_lb), coef);
        f |= lane_embed(store, var, dom, Itv(min(p, q), max(p, q)));
      }
      return f;
    }
    switch(kind) {
      case PC_EQ: f |= lane_embed(store, var, dom, pd); break;
      case PC_NEQ:
        if(len == 2) { if(pd.lb == pd.ub) f |= lane_embed(store, var, dom, itv_shave(dom, pd.lb)); }
        else f |= lane_embed(store, var, dom, itv_shave(dom, rhs));
        break;
      case PC_CLAUSE: {
        const int n_open = __popc(open);
        if((n_open == 1 && !refuted) || (n_open == 0 && lane == last)) f |= lane_embed(store, var, dom, coef < 0 ? Itv(0, 0) : Itv(1, 1));
        break;
      }
      case PC_ABS_EQ:
        if(pos == 0) f |= lane_embed(store, var, dom, fjoin(pd, Itv(b_neg(pd.ub), b_neg(pd.lb))));   // x <- hull(y, -y)
        else {                                                                                       // y <- |x|
          Itv ax = pd;
          if(!pd.is_bot()) {
            if(pd.lb >= 0) ax = pd;
            else if(pd.ub <= 0) ax = Itv(b_neg(pd.ub), b_neg(pd.lb));
            else ax = Itv(0, max(b_neg(pd.lb), pd.ub));
          }
          f |= lane_embed(store, var, dom, ax);
        }
        break;
      default: break;
    }
    return f;
  }
}

// One tile of one sweep under LPC_MODE_AUTO: skipped when `was` (what its lanes loaded at its last evaluation) equals what
// they load now (see the kernel).
template <bool BITS, class Acc>
__device__ __forceinline__ int tile_visit(const PcTableDev& t, Acc& acc, int2* store, long long tile, int lane, const int4 L,
                                          const int2 raw, int2* seen, bool have_was, const int2 was, unsigned& nprops) {
  if(seen != nullptr) {
    if(have_was && __all_sync(0xffffffffu, was.x == raw.x && was.y == raw.y)) return 0;
    __stcg(seen + tile * 32 + lane, raw);
    const int meta = L.z;
    nprops += __popc(__ballot_sync(0xffffffffu, PC_META_KIND(meta) != 0 && (meta & 31) == lane));
  }
  return tile_step<BITS>(t, acc, store, tile, lane, L, raw);
}

// The schedule (lpc_fixpoint_opts.mode), never the result:
//  * LPC_MODE_SWEEP (`seen` == nullptr): every sweep evaluates every tile.
//  * LPC_MODE_AUTO: `seen` holds the cells each lane of a tile loaded at the tile's last evaluation. A tile whose 32 cells
//    are what they were then is skipped, and exactly so: its evaluation is a function of those cells (cells only shrink,
//    so whatever it re-read in between had the same value), its result has been joined into the store already, and had it
//    moved one of its own cells that cell would differ now. A sweep that skips every tile moves nothing and ends the loop,
//    as in the dense schedule, on a store every propagator has been evaluated on: the same greatest common fixpoint. A
//    skipped tile costs two coalesced loads, a gather and a vote instead of ~300 instructions (config 3: 0.37 of the dense
//    schedule's propagator evaluations, 73 against 91 us; a sweep of tiles at rest still costs 8 us of L2 traffic).
//    Propagators outside the tiles (more than 32 lanes, trees) run every sweep.
//    Tried on top and dropped: flag bytes per tile, set through a var -> tiles index by whoever tightens a variable, so
//    that a sweep of tiles at rest reads one byte per tile. The late sweeps fall to 2-6 us, but walking the index inside
//    the divergent join costs more than it saves while many variables still move (first sweep of config 3: 42 against
//    32 us; of config 5, 13 tiles per variable: 174 against 35 us): tools/pc_sweep_probe.py, profiles/r02_pc_skip.txt.
template <bool BITS, int PC_UNROLL, int PC_MIN_BLOCKS>
__global__ void __launch_bounds__(PC_TPB, PC_MIN_BLOCKS) k_pc_fixpoint(PcTableDev t, int2* store, FixCtl* ctl, int max_sweeps,
                                                        int stop_on_bot, int2* seen) {
  typedef typename PcAcc<BITS>::type Acc;
  __shared__ unsigned long long s_vote;
  const int tid = threadIdx.x;
  const long long gtid = blockIdx.x * (long long)PC_TPB + tid;
  const long long gthreads = (long long)gridDim.x * PC_TPB;
  const int lane = tid & 31;
  const long long gwarp = gtid >> 5, gwarps = gthreads >> 5;
  int nbar = 0;
  bool bot;
  {   // a store that is already at bot stops before the first sweep
    int f = 0;
    for(long long i = gtid; i < t.nvars; i += gthreads) { int2 v = __ldcg(&store[i]); f |= BITS ? (v.x | v.y) == 0 : v.x > v.y; }
    bot = grid_vote_barrier(ctl->bar, nbar++, false, f != 0, &s_vote).bot;
  }
  int sweeps = 0;
  bool any_changed = false;
  bool done = (bot && stop_on_bot) || t.n == 0;
  unsigned nprops = 0;   // propagators this warp evaluated (lane-uniform)
  while(!done) {
    Acc acc = make_acc(store, (Acc*)nullptr);
    int f = 0;
    const bool hw = seen != nullptr && sweeps > 0;
    if constexpr(PC_UNROLL == 0) {
      // two-stage software pipeline: while tile i is evaluated, the domains of tile i + 1 and the records of tile i + 2
      // are in flight, so a warp waits for memory once per sweep instead of twice per tile
      const int4 Z = make_int4(0, 0, 0, 0);
      long long cur = gwarp;
      int4 L0 = cur < t.n_tiles ? __ldg(&t.tiles[cur * 32 + lane]) : Z;
      int4 L1 = cur + gwarps < t.n_tiles ? __ldg(&t.tiles[(cur + gwarps) * 32 + lane]) : Z;
      int2 r0 = PC_META_KIND(L0.z) ? __ldcg(&store[L0.y]) : make_int2(0, 0);
      while(cur < t.n_tiles) {
        const int4 L2 = cur + 2 * gwarps < t.n_tiles ? __ldg(&t.tiles[(cur + 2 * gwarps) * 32 + lane]) : Z;
        const int2 r1 = PC_META_KIND(L1.z) ? __ldcg(&store[L1.y]) : make_int2(0, 0);
        const int2 was = hw ? __ldcg(seen + cur * 32 + lane) : make_int2(0, 0);
        f |= tile_visit<BITS>(t, acc, store, cur, lane, L0, r0, seen, hw, was, nprops);
        L0 = L1; L1 = L2; r0 = r1;
        cur += gwarps;
      }
    }
    else
    for(long long base = gwarp * PC_UNROLL; base < t.n_tiles; base += gwarps * PC_UNROLL) {
      int4 L[PC_UNROLL ? PC_UNROLL : 1];
      int2 raw[PC_UNROLL ? PC_UNROLL : 1], was[PC_UNROLL ? PC_UNROLL : 1];
#pragma unroll
      for(int u = 0; u < PC_UNROLL; ++u) {
        L[u] = base + u < t.n_tiles ? __ldg(&t.tiles[(base + u) * 32 + lane]) : make_int4(0, 0, 0, 0);
        was[u] = hw && base + u < t.n_tiles ? __ldcg(seen + (base + u) * 32 + lane) : make_int2(0, 0);
      }
#pragma unroll
      for(int u = 0; u < PC_UNROLL; ++u) raw[u] = PC_META_KIND(L[u].z) ? __ldcg(&store[L[u].y]) : make_int2(0, 0);
#pragma unroll
      for(int u = 0; u < PC_UNROLL; ++u)
        if(base + u < t.n_tiles) f |= tile_visit<BITS>(t, acc, store, base + u, lane, L[u], raw[u], seen, hw, was[u], nprops);
    }
    for(long long b = gtid; b < t.n_big; b += gthreads) {
      const int4 h = t.hdr[t.big[b]];
      f |= pc_step<BITS>(acc, h, t.terms + h.y);
    }
    if(acc.seen_bot) f |= 2;
    f |= acc.touched & 1;
    const GridVote v = grid_vote_barrier(ctl->bar, nbar++, f & 1, f & 2, &s_vote);
    ++sweeps;
    bot |= v.bot;
    any_changed |= v.changed;
    if(!v.changed || (bot && stop_on_bot) || (max_sweeps && sweeps >= max_sweeps)) done = true;
  }
  if(blockIdx.x == 0 && tid == 0) {
    ctl->sweeps = sweeps;
    ctl->dense_sweeps = sweeps;
    ctl->has_changed = any_changed;
    ctl->is_bot = bot;
  }
  // deductions: every propagator every sweep (dense), or the propagators of the tiles actually evaluated + the untiled ones
  if(seen == nullptr) { if(blockIdx.x == 0 && tid == 0) ctl->deductions = (unsigned long long)sweeps * (unsigned long long)t.n; }
  else {
    if(lane == 0 && nprops) atomicAdd(&ctl->deductions, (unsigned long long)nprops);
    if(blockIdx.x == 0 && tid == 0) atomicAdd(&ctl->deductions, (unsigned long long)sweeps * (unsigned long long)t.n_big);
  }
}

template <bool BITS>
__global__ void k_pc_deduce_one(PcTableDev t, int2* store, long long i, int* out) {
  typedef typename PcAcc<BITS>::type Acc;
  Acc acc = make_acc(store, (Acc*)nullptr);
  const int4 h = t.hdr[i];
  out[0] = pc_step<BITS>(acc, h, t.terms + h.y) & 1;
}

template <bool BITS>
__global__ void k_pc_ask_all(PcTableDev t, const int2* store, unsigned long long* count, uint8_t* bits) {
  typedef typename PcAcc<BITS>::type Acc;
  unsigned cnt = 0;
  Acc acc = make_acc(const_cast<int2*>(store), (Acc*)nullptr);
  for(long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < t.n; p += (long long)gridDim.x * blockDim.x) {
    const int4 h = t.hdr[p];
    bool e;
    if constexpr(BITS) e = pc_ask_bits(acc, h, t.terms + h.y);
    else e = pc_ask(acc, h, t.terms + h.y);
    if(bits) bits[p] = e;
    cnt += e;
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if((threadIdx.x & 31) == 0 && cnt) atomicAdd(count, (unsigned long long)cnt);
}

} // namespace lpc

using namespace lpc;

// The device arrays of one view of a table.
struct PcArrays {
  PcTableDev dev{};
  void* d_hdr = nullptr;
  void* d_terms = nullptr;
  void* d_tiles = nullptr;
  void* d_tile_prop0 = nullptr;
  void* d_big = nullptr;
  void release() {
    cudaFree(d_hdr); cudaFree(d_terms); cudaFree(d_tiles); cudaFree(d_tile_prop0); cudaFree(d_big);
    d_hdr = d_terms = d_tiles = d_tile_prop0 = d_big = nullptr;
  }
};

struct lpc_pc_table {
  PcTableDev dev{};                 // the table as an interval store sees it (= main.dev)
  PcArrays main;
  // the table as an NBitset store sees it: linear kinds rewritten as formula streams (pc_bits_view). Built on the first
  // bitset call of a table that holds a linear kind; tables without one use `main` on both stores.
  PcArrays bits;
  bool bits_built = false;
  std::mutex bits_mu;               // a table is shared by the stores that run on it: the lazy build is serialised
  bool has_linear = false;          // a flat linear kind is present
  std::vector<int> h_props;         // the caller's arrays, kept for the bitset view (5 ints per propagator)
  std::vector<int2> h_terms;
  int sm_count = 0;
  int blocks_per_sm[2] = {0, 0};   // [interval store, bitset store]
  int variant = 0;
  lpc_store* host_store = nullptr;
};

static const void* pc_kernel(bool bits, int variant) {
  static const void* k[2][PC_NVAR] = {
    {(const void*)k_pc_fixpoint<false, 1, 4>, (const void*)k_pc_fixpoint<false, 2, 4>, (const void*)k_pc_fixpoint<false, 0, 3>,
     (const void*)k_pc_fixpoint<false, 0, 4>, (const void*)k_pc_fixpoint<false, 0, 2>},
    {(const void*)k_pc_fixpoint<true, 1, 4>, (const void*)k_pc_fixpoint<true, 2, 4>, (const void*)k_pc_fixpoint<true, 0, 3>,
     (const void*)k_pc_fixpoint<true, 0, 4>, (const void*)k_pc_fixpoint<true, 0, 2>}};
  return k[bits][variant];
}

// Check + upload one view: headers, terms, lane tiles (whole propagators packed into 32-lane tiles, in table order; see the
// header of this file), the list of untiled propagators.
static int pc_build_arrays(const lpc_pc_prop* props, int64_t n_props, const lpc_pc_term* terms, int64_t n_terms, int32_t nvars,
                           PcArrays* A, bool* has_linear_out) {
  std::vector<int4> hdr((size_t)std::max<int64_t>(n_props, 1));
  bool has_linear = false;
  for(int64_t i = 0; i < n_props; ++i) {
    const lpc_pc_prop& p = props[i];
    if(p.kind < LPC_PC_LIN_LE || p.kind > LPC_PC_TREE) { set_error("lpc_pc_table_create: propagator %lld has unsupported kind %d", (long long)i, p.kind); return LPC_ERR_UNSUPPORTED; }
    if(p.n_terms < 1 || p.n_terms >= (1 << 23) || p.first_term < 0 || (int64_t)p.first_term + p.n_terms > n_terms) { set_error("lpc_pc_table_create: propagator %lld has a bad term range", (long long)i); return LPC_ERR_INVALID; }
    if(p.kind == LPC_PC_TREE) {   // the term slots hold a prefix-encoded formula (pc_tree.cuh): checked here, walked on the device
      if(tree_check(reinterpret_cast<const int*>(terms + p.first_term), 2 * p.n_terms, nvars) < 0) {
        set_error("lpc_pc_table_create: propagator %lld is not a well-formed formula stream within the depth limits (terms %d, connectives %d)",
                  (long long)i, PC_TREE_TERM_DEPTH, PC_TREE_FORM_DEPTH);
        return LPC_ERR_UNSUPPORTED;
      }
      hdr[i] = make_int4(p.kind | (p.n_terms << 8), p.first_term, p.rhs, p.bvar);
      continue;
    }
    const bool two = p.kind == LPC_PC_EQ || p.kind == LPC_PC_ABS_EQ;
    if((two && p.n_terms != 2) || (p.kind == LPC_PC_NEQ && p.n_terms > 2)) { set_error("lpc_pc_table_create: propagator %lld has the wrong arity for its kind", (long long)i); return LPC_ERR_INVALID; }
    const bool lin_kind = pc_is_linear(p.kind);
    if((p.kind == LPC_PC_REIF_LIN_LE || p.kind == LPC_PC_LIN_EQ_VAR) && (p.bvar < 0 || p.bvar >= nvars)) { set_error("lpc_pc_table_create: propagator %lld has a bad reification / result variable", (long long)i); return LPC_ERR_INVALID; }
    for(int k = 0; k < p.n_terms; ++k) {
      const lpc_pc_term& t = terms[p.first_term + k];
      if(t.var < 0 || t.var >= nvars) { set_error("lpc_pc_table_create: propagator %lld has a variable out of range", (long long)i); return LPC_ERR_INVALID; }
      if((lin_kind || p.kind == LPC_PC_CLAUSE) && t.coef == 0) { set_error("lpc_pc_table_create: propagator %lld has a zero coefficient", (long long)i); return LPC_ERR_INVALID; }
    }
    hdr[i] = make_int4(p.kind | (p.n_terms << 8), p.first_term, p.rhs, p.bvar);
    has_linear |= lin_kind;
  }
  if(has_linear_out) *has_linear_out = has_linear;
  LPC_CUDA(cudaMalloc(&A->d_hdr, hdr.size() * sizeof(int4)));
  LPC_CUDA(cudaMalloc(&A->d_terms, std::max<size_t>((size_t)n_terms * sizeof(int2), 16)));
  if(n_props) LPC_CUDA(cudaMemcpy(A->d_hdr, hdr.data(), (size_t)n_props * sizeof(int4), cudaMemcpyHostToDevice));
  if(n_terms) LPC_CUDA(cudaMemcpy(A->d_terms, terms, (size_t)n_terms * sizeof(int2), cudaMemcpyHostToDevice));
  std::vector<int4> tiles;
  std::vector<int> prop0, big;
  int cur = 32;   // lanes used in the open tile (32 = none open)
  for(int64_t i = 0; i < n_props; ++i) {
    const lpc_pc_prop& p = props[i];
    const int lanes = p.n_terms + (pc_has_extra_lane(p.kind) ? 1 : 0);
    const bool lin_kind = pc_is_linear(p.kind);
    if(p.kind == LPC_PC_TREE || lanes > 32 || (lin_kind && (p.rhs >= (1 << 30) || p.rhs <= -(1 << 30)))) {   // see tile_step: keeps the tile arithmetic in int32
      big.push_back((int)i);
      cur = 32;   // close the tile so that tile_prop0 + rank stays a contiguous propagator range
      continue;
    }
    if(cur + lanes > 32) {
      tiles.resize(tiles.size() + 32, make_int4(0, 0, 0, 0));
      prop0.push_back((int)i);
      cur = 0;
    }
    int4* lane = tiles.data() + tiles.size() - 32 + cur;
    for(int k = 0; k < p.n_terms; ++k) {
      const lpc_pc_term& tm = terms[p.first_term + k];
      lane[k] = make_int4(tm.coef, tm.var, PC_META(cur, lanes, p.kind, 0), p.rhs);
    }
    if(lanes > p.n_terms) lane[p.n_terms] = make_int4(0, p.bvar, PC_META(cur, lanes, p.kind, 1), p.rhs);
    cur += lanes;
  }
  A->dev.n_tiles = (long long)prop0.size();
  A->dev.n_big = (int)big.size();
  LPC_CUDA(cudaMalloc(&A->d_tiles, std::max<size_t>(tiles.size() * sizeof(int4), 16)));
  LPC_CUDA(cudaMalloc(&A->d_tile_prop0, std::max<size_t>(prop0.size() * sizeof(int), 16)));
  LPC_CUDA(cudaMalloc(&A->d_big, std::max<size_t>(big.size() * sizeof(int), 16)));
  if(!tiles.empty()) LPC_CUDA(cudaMemcpy(A->d_tiles, tiles.data(), tiles.size() * sizeof(int4), cudaMemcpyHostToDevice));
  if(!prop0.empty()) LPC_CUDA(cudaMemcpy(A->d_tile_prop0, prop0.data(), prop0.size() * sizeof(int), cudaMemcpyHostToDevice));
  if(!big.empty()) LPC_CUDA(cudaMemcpy(A->d_big, big.data(), big.size() * sizeof(int), cudaMemcpyHostToDevice));
  A->dev.tiles = (const int4*)A->d_tiles; A->dev.tile_prop0 = (const int*)A->d_tile_prop0; A->dev.big = (const int*)A->d_big;
  A->dev.hdr = (const int4*)A->d_hdr; A->dev.terms = (const int2*)A->d_terms;
  A->dev.n = n_props; A->dev.n_terms = n_terms; A->dev.nvars = nvars;
  return LPC_OK;
}

// The view a store of the given universe runs on (the bitset view is built on first use).
static int pc_view(const lpc_pc_table* tc, bool bits, const PcTableDev** out) {
  lpc_pc_table* t = const_cast<lpc_pc_table*>(tc);
  if(!bits || !t->has_linear) { *out = &t->main.dev; return LPC_OK; }
  std::lock_guard<std::mutex> lock(t->bits_mu);
  if(!t->bits_built) {
    std::vector<int> vp;
    std::vector<int2> vt;
    pc_bits_view(t->h_props.data(), (long long)t->h_props.size() / 5, t->h_terms.data(), (long long)t->h_terms.size(), vp, vt);
    static_assert(sizeof(lpc_pc_prop) == 5 * sizeof(int) && sizeof(lpc_pc_term) == sizeof(int2), "flat layouts");
    int rc = pc_build_arrays(reinterpret_cast<const lpc_pc_prop*>(vp.data()), (int64_t)vp.size() / 5,
                             reinterpret_cast<const lpc_pc_term*>(vt.data()), (int64_t)vt.size(), t->dev.nvars, &t->bits, nullptr);
    if(rc) { t->bits.release(); return rc; }
    t->bits_built = true;
  }
  *out = &t->bits.dev;
  return LPC_OK;
}

extern "C" {

int lpc_pc_table_create(const lpc_pc_prop* props, int64_t n_props, const lpc_pc_term* terms, int64_t n_terms,
                        int32_t nvars, lpc_pc_table** out) {
  LPC_REQUIRE(out != nullptr, "null out");
  LPC_REQUIRE(n_props >= 0 && n_terms >= 0 && nvars >= 0, "negative size");
  LPC_REQUIRE((n_props == 0 || props) && (n_terms == 0 || terms), "null array");
  int cnt = 0;
  lpc_device_count(&cnt);
  if(cnt == 0) { set_error("no CUDA device: this library has no CPU path"); return LPC_ERR_NO_DEVICE; }
  lpc_pc_table* t = new lpc_pc_table();
  auto body = [&]() -> int {
    int rc = pc_build_arrays(props, n_props, terms, n_terms, nvars, &t->main, &t->has_linear);
    if(rc) return rc;
    t->dev = t->main.dev;
    if(t->has_linear) {   // kept for the bitset view
      t->h_props.assign(reinterpret_cast<const int*>(props), reinterpret_cast<const int*>(props) + 5 * n_props);
      t->h_terms.assign(reinterpret_cast<const int2*>(terms), reinterpret_cast<const int2*>(terms) + n_terms);
    }
    int dev = 0;
    LPC_CUDA(cudaGetDevice(&dev));
    LPC_CUDA(cudaDeviceGetAttribute(&t->sm_count, cudaDevAttrMultiProcessorCount, dev));
    t->variant = PC_VAR_DEFAULT;
    if(const char* e = getenv("LPC_PC_VARIANT")) { int v = atoi(e); if(v >= 0 && v < PC_NVAR) t->variant = v; }
    LPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&t->blocks_per_sm[0], pc_kernel(false, t->variant), PC_TPB, 0));
    LPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&t->blocks_per_sm[1], pc_kernel(true, t->variant), PC_TPB, 0));
    return LPC_OK;
  };
  const int rc = body();
  if(rc) { lpc_pc_table_destroy(t); return rc; }
  *out = t;
  return LPC_OK;
}

int lpc_pc_table_destroy(lpc_pc_table* t) {
  if(!t) return LPC_OK;
  t->main.release();
  t->bits.release();
  if(t->host_store) lpc_store_destroy(t->host_store);
  delete t;
  return LPC_OK;
}

int64_t lpc_pc_table_size(const lpc_pc_table* t) { return t ? (int64_t)t->dev.n : 0; }
int64_t lpc_pc_table_terms(const lpc_pc_table* t) { return t ? (int64_t)t->dev.n_terms : 0; }

static int pc_fixpoint_async(const lpc_pc_table* t, lpc_store* s, const lpc_fixpoint_opts* o, bool bits) {
  LPC_REQUIRE(t && s, "null argument");
  LPC_REQUIRE(s->nvars >= t->dev.nvars, "store smaller than the table's variable range");
  LPC_REQUIRE(t->blocks_per_sm[bits] > 0, "kernel does not fit on an SM");
  const PcTableDev* view = nullptr;
  if(int rc = pc_view(t, bits, &view)) return rc;
  lpc_fixpoint_opts def;
  if(!o) { lpc_fixpoint_default_opts(&def); o = &def; }
  cudaStream_t st = (cudaStream_t)o->stream;
  int grid = t->sm_count * t->blocks_per_sm[bits];
  long long want = std::max<long long>({1, (view->n_tiles * 32 + PC_TPB - 1) / PC_TPB, ((long long)view->n_big + PC_TPB - 1) / PC_TPB,
                                        ((long long)s->nvars + 16 * PC_TPB - 1) / (16 * PC_TPB)});   // the last: the prologue's store scan
  if(want < grid) grid = (int)want;
  int2* seen = nullptr;
  // per tile: the 32 cells last seen. Only where a tile is expensive to evaluate (sums: scan, divisions): the tiles of =, !=,
  // clauses and abs cost about what the comparison costs (config 5: 0.119 against 0.114 ms), so they always run
  if(o->mode != LPC_MODE_SWEEP && view->n_tiles > 0 && t->has_linear && !bits) {
    const long long cells = view->n_tiles * 32;
    if(s->pc_seen_cap < cells) {
      cudaFree(s->d_pc_seen);
      s->d_pc_seen = nullptr; s->pc_seen_cap = 0;
      LPC_CUDA(cudaMalloc((void**)&s->d_pc_seen, (size_t)cells * sizeof(int2)));
      s->pc_seen_cap = cells;
    }
    seen = s->d_pc_seen;
  }
  LPC_CUDA(cudaEventRecord(s->ev0, st));
  LPC_CUDA(cudaMemsetAsync(s->d_ctl, 0, sizeof(FixCtl), st));
  PcTableDev td = *view;
  int2* store = s->d;
  FixCtl* ctl = s->d_ctl;
  int max_sweeps = o->max_sweeps, stop = o->stop_on_bot;
  void* args[] = {&td, &store, &ctl, &max_sweeps, &stop, &seen};
  LPC_CUDA(cudaLaunchCooperativeKernel(pc_kernel(bits, t->variant), dim3(grid), dim3(PC_TPB), args, 0, st));
  g_launches++;
  LPC_CUDA(cudaEventRecord(s->ev1, st));
  LPC_CUDA(cudaMemcpyAsync(s->h_ctl, s->d_ctl, sizeof(FixCtl), cudaMemcpyDeviceToHost, st));
  s->last_stream = st;
  s->pending = true;
  return LPC_OK;
}

static int pc_fixpoint_host(const lpc_pc_table* t, void* cells, const lpc_fixpoint_opts* o, lpc_fixpoint_result* r, bool bits) {
  LPC_REQUIRE(t && cells, "null argument");
  lpc_pc_table* tm = const_cast<lpc_pc_table*>(t);
  if(!tm->host_store) {
    int rc = lpc_store_create(t->dev.nvars, &tm->host_store);
    if(rc) return rc;
  }
  lpc_store* s = tm->host_store;
  cudaStream_t st = o ? (cudaStream_t)o->stream : nullptr;
  size_t bytes = (size_t)t->dev.nvars * 8;
  if(bytes) LPC_CUDA(cudaMemcpyAsync(s->d, cells, bytes, cudaMemcpyHostToDevice, st));
  int rc = pc_fixpoint_async(t, s, o, bits);
  if(rc) return rc;
  if(bytes) LPC_CUDA(cudaMemcpyAsync(cells, s->d, bytes, cudaMemcpyDeviceToHost, st));
  return lpc_fixpoint_collect(s, r);
}

static int pc_deduce_one(const lpc_pc_table* t, lpc_store* s, int64_t i, int* changed, bool bits) {
  LPC_REQUIRE(t && s, "null argument");
  LPC_REQUIRE(i >= 0 && i < t->dev.n, "propagator index out of range");
  LPC_REQUIRE(s->nvars >= t->dev.nvars, "store smaller than the table's variable range");
  const PcTableDev* view = nullptr;
  if(int rc = pc_view(t, bits, &view)) return rc;
  if(bits) k_pc_deduce_one<true><<<1, 1>>>(*view, s->d, i, &s->d_ctl->scratch[0]);
  else k_pc_deduce_one<false><<<1, 1>>>(*view, s->d, i, &s->d_ctl->scratch[0]);
  g_launches++;
  LPC_CUDA(cudaGetLastError());
  int c = 0;
  LPC_CUDA(cudaMemcpy(&c, &s->d_ctl->scratch[0], sizeof(int), cudaMemcpyDeviceToHost));
  if(changed) *changed = c;
  return LPC_OK;
}

static int pc_ask_all(const lpc_pc_table* t, const lpc_store* s, int64_t* n_entailed, uint8_t* bits_out, bool bits) {
  LPC_REQUIRE(t && s, "null argument");
  LPC_REQUIRE(s->nvars >= t->dev.nvars, "store smaller than the table's variable range");
  const PcTableDev* view = nullptr;
  if(int rc = pc_view(t, bits, &view)) return rc;
  unsigned long long* d_cnt = nullptr;
  uint8_t* d_bits = nullptr;
  LPC_CUDA(cudaMalloc((void**)&d_cnt, 8));
  LPC_CUDA(cudaMemset(d_cnt, 0, 8));
  if(bits_out && t->dev.n) LPC_CUDA(cudaMalloc((void**)&d_bits, t->dev.n));
  if(t->dev.n) {
    int blocks = (int)std::min<long long>(ceil_div(t->dev.n, 256), 148 * 8);
    if(bits) k_pc_ask_all<true><<<blocks, 256>>>(*view, s->d, d_cnt, d_bits);
    else k_pc_ask_all<false><<<blocks, 256>>>(*view, s->d, d_cnt, d_bits);
    g_launches++;
    LPC_CUDA(cudaGetLastError());
  }
  unsigned long long c = 0;
  LPC_CUDA(cudaMemcpy(&c, d_cnt, 8, cudaMemcpyDeviceToHost));
  if(bits_out && t->dev.n) LPC_CUDA(cudaMemcpy(bits_out, d_bits, t->dev.n, cudaMemcpyDeviceToHost));
  cudaFree(d_cnt);
  cudaFree(d_bits);
  if(n_entailed) *n_entailed = (int64_t)c;
  return LPC_OK;
}

int lpc_pc_fixpoint(const lpc_pc_table* t, lpc_store* s, const lpc_fixpoint_opts* o, lpc_fixpoint_result* r) {
  int rc = pc_fixpoint_async(t, s, o, false);
  if(rc) return rc;
  return lpc_fixpoint_collect(s, r);
}
int lpc_pc_fixpoint_host(const lpc_pc_table* t, int32_t* lbub, const lpc_fixpoint_opts* o, lpc_fixpoint_result* r) {
  return pc_fixpoint_host(t, lbub, o, r, false);
}
int lpc_pc_deduce_one(const lpc_pc_table* t, lpc_store* s, int64_t i, int* changed) { return pc_deduce_one(t, s, i, changed, false); }
int lpc_pc_ask_all(const lpc_pc_table* t, const lpc_store* s, int64_t* n_entailed, uint8_t* bits) {
  return pc_ask_all(t, s, n_entailed, bits, false);
}

/* ---- bitset stores ------------------------------------------------------------------------------------ */
int lpc_store_write_bits(lpc_store* s, int32_t first, int32_t n, const uint64_t* cells) {
  return lpc_store_write(s, first, n, reinterpret_cast<const int32_t*>(cells));
}
int lpc_store_read_bits(const lpc_store* s, int32_t first, int32_t n, uint64_t* cells) {
  return lpc_store_read(s, first, n, reinterpret_cast<int32_t*>(cells));
}
int lpc_pc_fixpoint_bits(const lpc_pc_table* t, lpc_store* s, const lpc_fixpoint_opts* o, lpc_fixpoint_result* r) {
  int rc = pc_fixpoint_async(t, s, o, true);
  if(rc) return rc;
  return lpc_fixpoint_collect(s, r);
}
int lpc_pc_fixpoint_bits_host(const lpc_pc_table* t, uint64_t* cells, const lpc_fixpoint_opts* o, lpc_fixpoint_result* r) {
  return pc_fixpoint_host(t, cells, o, r, true);
}
int lpc_pc_deduce_one_bits(const lpc_pc_table* t, lpc_store* s, int64_t i, int* changed) { return pc_deduce_one(t, s, i, changed, true); }
int lpc_pc_ask_all_bits(const lpc_pc_table* t, const lpc_store* s, int64_t* n_entailed, uint8_t* bits) {
  return pc_ask_all(t, s, n_entailed, bits, true);
}
uint64_t lpc_nbit_range(int32_t lb, int32_t ub) {
  if(lb > ub) return 0;
  const int from = lb < 0 ? 0 : (lb >= 62 ? 63 : lb + 1), to = ub < 0 ? 0 : (ub >= 62 ? 63 : ub + 1);
  return (~0ull << from) & (~0ull >> (63 - to));
}

} // extern "C"
