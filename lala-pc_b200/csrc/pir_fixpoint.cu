// pir_fixpoint.cu — the single-store fixpoint: a persistent cooperative kernel for sm_100a.
//
// Replaces the loop  GaussSeidelIteration{}.fixpoint(pir.num_deductions(), [&](size_t i){ return pir.deduce(i); })
// (call sites tests/pir_test.cpp:60-62, 82-86) with chaotic parallel iteration. Because every propagator is
// monotone, the least fixpoint does not depend on the schedule; only sweep and deduction counts do.
//
// One launch = one fixpoint:
//   dense sweeps   the table is cut into its opcode segments (it is sorted by (op, y, x, z), pir.hpp:343-347) and
//                  every resident block owns the same contiguous fraction of EVERY segment: contiguous so that the
//                  sort order turns into L1 locality of the gathers, per segment so that all blocks see the same
//                  operator mix and reach the grid barrier together. A thread fetches RPT records with vector loads
//                  (RPT = 4: one 32-bit + three 128-bit loads), gathers their intervals, evaluates the rules in
//                  registers and joins tightened bounds with atomicMax / atomicMin (RED at L2).
//   worklist       once few variables change per sweep, the kernel switches to change-driven iterations: changed
//                  variables enqueue their incident propagators (var -> records CSR, warp-cooperative, de-duplicated
//                  by an iteration stamp) and only those are re-run.
//   termination    a grid barrier per iteration; the barrier's fence invalidates L1, so the last (quiescent)
//                  iteration reads the final store. has_changed / bot flags are formed by __syncthreads_or and one
//                  atomic per block.
#include "lpc_internal.cuh"
#include "grid_barrier.cuh"

#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace cg = cooperative_groups;

namespace lpc {

constexpr int TPB = 256;
constexpr int MAX_SEG = 16;

// Opcode segments of the padded table, in quads (4 records). nseg == 0 never happens (an empty table has one
// all-padding segment).
struct SegTable {
  int nseg;
  int q[MAX_SEG + 1];
};

// Join the new domain into the store. Returns bit0 = changed, bit1 = became empty.
template <bool TRACK>
__device__ __forceinline__ int commit(int2* p, int2 old, const Itv& nw, int* vmark, int v, int mark) {
  int f = 0;
  if(nw.lb > old.x) { atomicMax(&p->x, nw.lb); f = 1; }
  if(nw.ub < old.y) { atomicMin(&p->y, nw.ub); f = 1; }
  if(f) {
    if(nw.lb > nw.ub) f |= 2;
    if(TRACK) vmark[v] = mark;
  }
  return f;
}

// One propagator evaluation: rules in registers, then (rarely) the joins. The common case - nothing tightens, no
// operand empty - is nine compares folded into one predicate and a single not-taken branch.
template <bool HAS_DIV, bool TRACK>
__device__ __forceinline__ int run_record(int op, int xi, int yi, int zi, int2 a, int2 b, int2 c, int2* store,
                                          int* vmark, int mark) {
  Itv r1(a.x, a.y), r2(b.x, b.y), r3(c.x, c.y);
  deduce_regs<HAS_DIV>(op, r1, r2, r3);
  const bool slow = (r1.lb > a.x) | (r1.ub < a.y) | (r2.lb > b.x) | (r2.ub < b.y) | (r3.lb > c.x) | (r3.ub < c.y)
                  | (a.x > a.y) | (b.x > b.y) | (c.x > c.y);
  int f = 0;
  if(slow) {
    f = ((a.x > a.y) | (b.x > b.y) | (c.x > c.y)) ? 2 : 0;
    f |= commit<TRACK>(store + xi, a, r1, vmark, xi, mark);
    f |= commit<TRACK>(store + yi, b, r2, vmark, yi, mark);
    f |= commit<TRACK>(store + zi, c, r3, vmark, zi, mark);
  }
  return f;
}

struct WlState {
  int* stamp;      // [n] iteration id at which a record was last enqueued
  int* queue[2];   // [n] each
  int* vmark;      // [nvars] sweep id at which a variable last changed
};

// Enqueue every propagator incident to variable v (warp-cooperative, warp-aggregated append).
__device__ __forceinline__ void expand_var(int v, int lane, int stampval, const TableDev& t, const WlState& w,
                                           int* qnext, int* qlen_next) {
  const int b = t.inc_off[v], e = t.inc_off[v + 1];
  for(int base = b; base < e; base += 32) {
    int j = base + lane;
    int r = j < e ? t.inc_idx[j] : -1;
    bool push = r >= 0 && atomicExch(&w.stamp[r], stampval) != stampval;
    unsigned m = __ballot_sync(0xffffffffu, push);
    if(m) {
      int leader = __ffs(m) - 1, pos = 0;
      if(lane == leader) pos = atomicAdd(qlen_next, __popc(m));
      pos = __shfl_sync(0xffffffffu, pos, leader);
      if(push) qnext[pos + __popc(m & ((1u << lane) - 1))] = r;
    }
  }
}

// RPT records of unit u: opcode bytes and x / y / z indices with one vector load each.
template <int RPT> struct Unit;
template <> struct Unit<4> {
  int op[4], x[4], y[4], z[4];
  __device__ __forceinline__ void load(const TableDev& t, int u) {
    const uchar4 o = reinterpret_cast<const uchar4*>(t.op)[u];
    const int4 X = reinterpret_cast<const int4*>(t.x)[u], Y = reinterpret_cast<const int4*>(t.y)[u],
               Z = reinterpret_cast<const int4*>(t.z)[u];
    op[0] = o.x; op[1] = o.y; op[2] = o.z; op[3] = o.w;
    x[0] = X.x; x[1] = X.y; x[2] = X.z; x[3] = X.w;
    y[0] = Y.x; y[1] = Y.y; y[2] = Y.z; y[3] = Y.w;
    z[0] = Z.x; z[1] = Z.y; z[2] = Z.z; z[3] = Z.w;
  }
};
template <> struct Unit<2> {
  int op[2], x[2], y[2], z[2];
  __device__ __forceinline__ void load(const TableDev& t, int u) {
    const uchar2 o = reinterpret_cast<const uchar2*>(t.op)[u];
    const int2 X = reinterpret_cast<const int2*>(t.x)[u], Y = reinterpret_cast<const int2*>(t.y)[u],
               Z = reinterpret_cast<const int2*>(t.z)[u];
    op[0] = o.x; op[1] = o.y;
    x[0] = X.x; x[1] = X.y; y[0] = Y.x; y[1] = Y.y; z[0] = Z.x; z[1] = Z.y;
  }
};
template <> struct Unit<1> {
  int op[1], x[1], y[1], z[1];
  __device__ __forceinline__ void load(const TableDev& t, int u) {
    op[0] = t.op[u]; x[0] = t.x[u]; y[0] = t.y[u]; z[0] = t.z[u];
  }
};

// switch_at = number of change events per sweep at or below which the worklist takes over
// (0: never, UINT_MAX: after the first sweep).
template <bool HAS_DIV, bool TRACK, int RPT, int MINB, bool PIPE>
__global__ void __launch_bounds__(TPB, MINB) k_pir_fixpoint(TableDev t, int2* store, SegTable seg, FixCtl* ctl,
                                                            WlState w, int max_sweeps, int stop_on_bot,
                                                            unsigned switch_at, int use_vote, int sm_order) {
  cg::grid_group grid = cg::this_grid();
  __shared__ unsigned long long s_vote;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const long long gtid = blockIdx.x * (long long)TPB + tid;
  const long long gthreads = (long long)gridDim.x * TPB;
  volatile int* vflags = ctl->flags;
  volatile int* vbot = &ctl->is_bot;
  volatile int* vqlen = ctl->q_len;   // 3 rotating queue-length words

  // ---- prologue: a store created with an empty variable is at bot before the first sweep ----
  const int sm_slot = sm_order ? sm_rank_arrive(ctl->sm_slots) : 0;
  {
    int f = 0;
    for(long long i = gtid; i < t.nvars; i += gthreads) { int2 v = store[i]; f |= v.x > v.y; }
    if(__syncthreads_or(f) && tid == 0) atomicOr(&ctl->is_bot, 1);
  }
  grid.sync();
  // which fraction of every segment this block sweeps: its rank in SM order (co-resident blocks get adjacent fractions)
  // sm_order 1: this block's own fraction; 2: the SM's blocks interleave over the SM's fraction (one moving window)
  long long bid = blockIdx.x, bspan = 1;
  int bslot = 0;
  if(sm_order) {
    const SmRank r = sm_rank_resolve(ctl->sm_slots, sm_slot);
    if(sm_order == 2 && !PIPE) { bid = r.below; bspan = r.here; bslot = r.slot; }
    else bid = r.below + r.slot;
  }
  int sweeps = 0, dense = 0;
  bool any_changed = false;
  bool bot = *vbot != 0;
  unsigned long long deductions = 0;
  bool done = (bot && stop_on_bot) || t.n == 0;
  bool worklist = false;
  constexpr int UPQ = 4 / RPT;   // units per quad

  // ---- dense sweeps ----
  while(!done) {
    const int slot = sweeps % 3;
    if(blockIdx.x == 0 && tid == 0) vflags[(sweeps + 1) % 3] = 0;
    const int mark = sweeps + 1;
    int f = 0, nchg = 0;
    for(int s = 0; s < seg.nseg; ++s) {
      const long long sq0 = seg.q[s], len = seg.q[s + 1] - sq0;
      const int u0 = (int)((sq0 + len * bid / gridDim.x) * UPQ);
      const int u1 = (int)((sq0 + len * (bid + bspan) / gridDim.x) * UPQ);
      if constexpr(PIPE) {
        // two-stage software pipeline: while unit i is evaluated, the bounds of unit i + 1 and the records of unit
        // i + 2 are in flight, so a thread waits for memory once per segment instead of twice per unit. Indices past
        // the end are clamped to the last unit of the range (a redundant load, never evaluated).
        int u = u0 + tid;
        if(u < u1) {
          const int ulast = u1 - 1;
          Unit<RPT> cur, nxt;
          cur.load(t, u);
          nxt.load(t, min(u + TPB, ulast));
          int2 a[RPT], b[RPT], c[RPT];
#pragma unroll
          for(int k = 0; k < RPT; ++k) { a[k] = store[cur.x[k]]; b[k] = store[cur.y[k]]; c[k] = store[cur.z[k]]; }
          while(true) {
            Unit<RPT> nn;
            nn.load(t, min(u + 2 * TPB, ulast));
            int2 a2[RPT], b2[RPT], c2[RPT];
#pragma unroll
            for(int k = 0; k < RPT; ++k) { a2[k] = store[nxt.x[k]]; b2[k] = store[nxt.y[k]]; c2[k] = store[nxt.z[k]]; }
#pragma unroll
            for(int k = 0; k < RPT; ++k) {
              const int g = run_record<HAS_DIV, TRACK>(cur.op[k], cur.x[k], cur.y[k], cur.z[k], a[k], b[k], c[k], store, w.vmark, mark);
              f |= g;
              nchg += g & 1;
            }
            u += TPB;
            if(u >= u1) break;
            cur = nxt; nxt = nn;
#pragma unroll
            for(int k = 0; k < RPT; ++k) { a[k] = a2[k]; b[k] = b2[k]; c[k] = c2[k]; }
          }
        }
      }
      else
      for(int u = u0 + bslot * TPB + tid; u < u1; u += (int)bspan * TPB) {
        Unit<RPT> r;
        r.load(t, u);
        int2 a[RPT], b[RPT], c[RPT];
#pragma unroll
        for(int k = 0; k < RPT; ++k) { a[k] = store[r.x[k]]; b[k] = store[r.y[k]]; c[k] = store[r.z[k]]; }
#pragma unroll
        for(int k = 0; k < RPT; ++k) {
          const int g = run_record<HAS_DIV, TRACK>(r.op[k], r.x[k], r.y[k], r.z[k], a[k], b[k], c[k], store, w.vmark, mark);
          f |= g;
          nchg += g & 1;
        }
      }
    }
    if(!TRACK && use_vote) {
      // the arrival at the grid barrier carries the block's votes (grid_barrier.cuh): one L2 round trip instead of four
      const GridVote v = grid_vote_barrier(ctl->bar, sweeps, f & 1, f & 2, &s_vote);
      ++sweeps; ++dense;
      deductions += (unsigned long long)t.n;
      bot |= v.bot;
      any_changed |= v.changed;
      if(!v.changed || (bot && stop_on_bot) || (max_sweeps && sweeps >= max_sweeps)) done = true;
    }
    else {
      // block-level flags: one atomic per block
      if(__syncthreads_or(f & 2) && tid == 0) atomicOr(&ctl->is_bot, 1);
      if(TRACK) {
        nchg = __reduce_add_sync(0xffffffffu, nchg);
        if(lane == 0 && nchg) atomicAdd(&ctl->flags[slot], nchg);
      }
      else if(__syncthreads_or(f & 1) && tid == 0) atomicOr(&ctl->flags[slot], 1);
      grid.sync();
      ++sweeps; ++dense;
      deductions += (unsigned long long)t.n;
      const unsigned c = (unsigned)vflags[slot];
      bot = *vbot != 0;
      any_changed |= c != 0;
      if(c == 0 || (bot && stop_on_bot) || (max_sweeps && sweeps >= max_sweeps)) done = true;
      else if(TRACK && c <= switch_at) { worklist = true; break; }
    }
  }

  if(TRACK && worklist && !done) {
    // ---- frontier expansion: variables marked in the last dense sweep enqueue their propagators ----
    const int mark = sweeps;        // marks written by the last dense sweep
    int it = sweeps + 1;            // stamp value == iteration id, strictly increasing
    // queue-length words rotate over 3 slots; slot (it % 3) is the one being filled for iteration `it`
    {
      const long long warps = gthreads >> 5, wid = gtid >> 5;
      for(long long base = wid * 32; base < t.nvars; base += warps * 32) {
        long long v = base + lane;
        bool hit = v < t.nvars && w.vmark[v] == mark;
        unsigned m = __ballot_sync(0xffffffffu, hit);
        while(m) {
          int src = __ffs(m) - 1;
          m &= m - 1;
          expand_var((int)(base + src), lane, it, t, w, w.queue[it & 1], (int*)&vqlen[it % 3]);
        }
      }
    }
    grid.sync();
    // ---- change-driven iterations ----
    while(true) {
      const int len = vqlen[it % 3];
      if(len == 0) break;
      if(blockIdx.x == 0 && tid == 0) vqlen[(it + 2) % 3] = 0;   // last read in iteration it-1
      const int* qcur = w.queue[it & 1];
      int* qnext = w.queue[(it + 1) & 1];
      int* qlen_next = (int*)&vqlen[(it + 1) % 3];
      int f = 0;
      const long long warps = gthreads >> 5, wid = gtid >> 5;
      for(long long base = wid * 32; base < len; base += warps * 32) {
        long long i = base + lane;
        int cv0 = -1, cv1 = -1, cv2 = -1;
        if(i < len) {
          const int r = qcur[i];
          const int op = t.op[r], xi = t.x[r], yi = t.y[r], zi = t.z[r];
          const int2 a = store[xi], b = store[yi], c = store[zi];
          Itv r1(a.x, a.y), r2(b.x, b.y), r3(c.x, c.y);
          if(r1.is_bot() | r2.is_bot() | r3.is_bot()) f |= 2;
          deduce_regs<HAS_DIV>(op, r1, r2, r3);
          int g0 = commit<false>(store + xi, a, r1, nullptr, 0, 0);
          int g1 = commit<false>(store + yi, b, r2, nullptr, 0, 0);
          int g2 = commit<false>(store + zi, c, r3, nullptr, 0, 0);
          if(g0 & 1) cv0 = xi;
          if(g1 & 1) cv1 = yi;
          if(g2 & 1) cv2 = zi;
          f |= g0 | g1 | g2;
        }
#pragma unroll
        for(int k = 0; k < 3; ++k) {
          int cv = k == 0 ? cv0 : k == 1 ? cv1 : cv2;
          unsigned m = __ballot_sync(0xffffffffu, cv >= 0);
          while(m) {
            int src = __ffs(m) - 1;
            m &= m - 1;
            int v = __shfl_sync(0xffffffffu, cv, src);
            expand_var(v, lane, it + 1, t, w, qnext, qlen_next);
          }
        }
      }
      if(__syncthreads_or(f & 2) && tid == 0) atomicOr(&ctl->is_bot, 1);
      grid.sync();
      ++sweeps;
      deductions += (unsigned long long)len;
      ++it;
      bot = *vbot != 0;
      if((bot && stop_on_bot) || (max_sweeps && sweeps >= max_sweeps)) break;
    }
  }

  if(blockIdx.x == 0 && tid == 0) {
    ctl->sweeps = sweeps;
    ctl->dense_sweeps = dense;
    ctl->has_changed = any_changed;
    ctl->is_bot = bot;
    ctl->deductions = deductions;
  }
}

// ---- single-step and entailment kernels ----------------------------------------------------------------------------
__global__ void k_deduce_one(TableDev t, int2* store, long long i, int* out) {
  const int op = t.op[i], xi = t.x[i], yi = t.y[i], zi = t.z[i];
  int2 a = store[xi], b = store[yi], c = store[zi];
  int f = run_record<true, false>(op, xi, yi, zi, a, b, c, store, nullptr, 0);
  out[0] = f & 1;
}

__global__ void k_ask_all(TableDev t, const int2* store, unsigned long long* count, uint8_t* bits) {
  unsigned cnt = 0;
  for(long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < t.n; i += (long long)gridDim.x * blockDim.x) {
    const int op = t.op[i];
    int2 a = store[t.x[i]], b = store[t.y[i]], c = store[t.z[i]];
    bool e = ask_regs(op, Itv(a.x, a.y), Itv(b.x, b.y), Itv(c.x, c.y));
    if(bits) bits[i] = e;
    cnt += e;
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if((threadIdx.x & 31) == 0 && cnt) atomicAdd(count, (unsigned long long)cnt);
}

typedef void (*fix_kernel_t)(TableDev, int2*, SegTable, FixCtl*, WlState, int, int, unsigned, int, int);

// The (records per thread, min blocks per SM) variants that are built; LPC_RPT / LPC_MINB select one for tuning.
struct Variant { int rpt, minb; fix_kernel_t k[2][2]; };
#define LPC_VARIANT(R, M, P) { R + 10 * P, M, { { k_pir_fixpoint<false, false, R, M, P>, k_pir_fixpoint<false, true, R, M, P> }, \
                                    { k_pir_fixpoint<true, false, R, M, P>, k_pir_fixpoint<true, true, R, M, P> } } }
// LPC_RPT = records per thread (+ 10 for the software-pipelined loop)
static const Variant kVariants[] = { LPC_VARIANT(4, 2, false), LPC_VARIANT(4, 3, false), LPC_VARIANT(2, 3, false), LPC_VARIANT(2, 4, false),
                                     LPC_VARIANT(1, 4, false), LPC_VARIANT(1, 4, true), LPC_VARIANT(1, 3, true), LPC_VARIANT(2, 3, true),
                                     LPC_VARIANT(2, 2, true) };
static const int kDefaultVariant = 2;   // RPT 2, 3 blocks / SM: fastest on config 2 (profiles/r01_variants.md)

static const Variant& pick_variant() {
  static int chosen = -1;
  if(chosen < 0) {
    chosen = kDefaultVariant;
    const char* r = getenv("LPC_RPT");
    const char* m = getenv("LPC_MINB");
    if(r && m) {
      for(size_t i = 0; i < sizeof(kVariants) / sizeof(kVariants[0]); ++i)
        if(kVariants[i].rpt == atoi(r) && kVariants[i].minb == atoi(m)) chosen = (int)i;
    }
  }
  return kVariants[chosen];
}

// Opcode segments of the padded table in quads. A table that is not sorted by opcode (more than MAX_SEG runs) is
// treated as one segment.
static SegTable build_segments(const lpc_table* t) {
  SegTable s;
  const long long nq = t->dev.n_pad / 4;
  auto dev_op = [&](long long i) -> int {
    if(i >= (long long)t->host.size()) return D_NOP;
    switch(t->host[i].op) {
      case LPC_ADD: return D_ADD; case LPC_MUL: return D_MUL; case LPC_MIN: return D_MIN; case LPC_MAX: return D_MAX;
      case LPC_TDIV: return D_TDIV; case LPC_FDIV: return D_FDIV; case LPC_CDIV: return D_CDIV;
      case LPC_EDIV: return D_EDIV; case LPC_EQ: return D_EQ; default: return D_LEQ;
    }
  };
  s.nseg = 0;
  s.q[0] = 0;
  int prev = dev_op(0);
  for(long long q = 1; q < nq; ++q) {
    int o = dev_op(q * 4);
    if(o != prev) {
      if(s.nseg + 1 >= MAX_SEG) { s.nseg = 0; break; }
      s.q[++s.nseg] = (int)q;
      prev = o;
    }
  }
  s.q[++s.nseg] = (int)nq;
  return s;
}

} // namespace lpc

using namespace lpc;

// worklist scratch lives with the store (one in-flight call per store handle)
static int get_scratch(lpc_store* s, const lpc_table* t, WlState* w) {
  long long n = std::max<long long>(t->dev.n, 1);
  if(s->wl_n < n || s->wl_nvars < s->nvars) {
    cudaFree(s->wl_stamp); cudaFree(s->wl_q0); cudaFree(s->wl_q1); cudaFree(s->wl_vmark);
    s->wl_stamp = s->wl_q0 = s->wl_q1 = s->wl_vmark = nullptr;
    s->wl_n = 0; s->wl_nvars = 0;
    LPC_CUDA(cudaMalloc((void**)&s->wl_stamp, n * 4));
    LPC_CUDA(cudaMalloc((void**)&s->wl_q0, n * 4));
    LPC_CUDA(cudaMalloc((void**)&s->wl_q1, n * 4));
    LPC_CUDA(cudaMalloc((void**)&s->wl_vmark, (size_t)std::max(1, s->nvars) * 4));
    s->wl_n = n; s->wl_nvars = s->nvars;
  }
  w->stamp = s->wl_stamp; w->queue[0] = s->wl_q0; w->queue[1] = s->wl_q1; w->vmark = s->wl_vmark;
  return LPC_OK;
}

extern "C" {

void lpc_fixpoint_default_opts(lpc_fixpoint_opts* o) {
  if(!o) return;
  memset(o, 0, sizeof(*o));
  o->mode = LPC_MODE_AUTO;
  o->max_sweeps = 0;
  o->stop_on_bot = 1;
}

int lpc_fixpoint_async(const lpc_table* tc, lpc_store* s, const lpc_fixpoint_opts* o) {
  LPC_REQUIRE(tc && s, "null argument");
  lpc_table* t = const_cast<lpc_table*>(tc);
  LPC_REQUIRE(s->nvars >= t->dev.nvars, "store smaller than the table's variable range");
  lpc_fixpoint_opts def;
  if(!o) { lpc_fixpoint_default_opts(&def); o = &def; }
  LPC_REQUIRE(o->mode >= LPC_MODE_AUTO && o->mode <= LPC_MODE_WORKLIST, "bad mode");
  cudaStream_t st = (cudaStream_t)o->stream;
  // LPC_MODE_AUTO is the change-driven kernel of pir_dirty.cu (dense sweeps that skip unflagged 64-record groups once
  // few groups change); LPC_MODE_WORKLIST keeps the record-granular queues of this file; LPC_MODE_SWEEP is purely dense.
  const bool track = o->mode == LPC_MODE_WORKLIST;
  const Variant& var = pick_variant();
  fix_kernel_t k = var.k[t->has_div ? 1 : 0][track ? 1 : 0];
  // launch plan: computed once per table (the segment scan is O(n) on the host, the occupancy query is a driver call)
  if(!t->plan_ready) {
    SegTable sg = build_segments(t);
    t->seg_n = sg.nseg;
    for(int i = 0; i <= sg.nseg; ++i) t->seg_q[i] = sg.q[i];
    const char* ce = getenv("LPC_CARVE");   // see pir_dirty.cu
    if(!ce || atoi(ce))
      for(int tr = 0; tr < 2; ++tr)
        LPC_CUDA(cudaFuncSetAttribute((const void*)var.k[t->has_div ? 1 : 0][tr], cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1));
    for(int tr = 0; tr < 2; ++tr)
      LPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&t->blocks_per_sm[tr], var.k[t->has_div ? 1 : 0][tr], TPB, 0));
    t->plan_ready = true;
  }
  if(!track) {   // a small network runs inside one thread-block cluster, replicas of the store in shared memory (pir_cluster.cu)
    int used = 0;
    int rc = lpc_cluster_fixpoint_launch(t, s, o, &used);
    if(rc || used) return rc;
  }
  // AUTO on a table with at most a couple of units per thread stays with the plain dense kernel: such a table is swept in
  // a few microseconds, never reaches the flagged phase, and only pays the change-driven kernel's bookkeeping
  // (config 1: 115 vs 107 us)
  const bool small_table = t->dev.n_pad / 2 <= 2LL * t->sm_count * 3 * TPB;
  if(o->mode == LPC_MODE_AUTO && !small_table) return lpc_dirty_fixpoint_launch(t, s, o);
  if(!track) {   // dense sweeps: shared-memory windows when the table has the locality for it (pir_window.cu)
    int used = 0;
    int rc = lpc_win_fixpoint_launch(t, s, o, &used);
    if(rc || used) return rc;
  }
  const int per_sm = t->blocks_per_sm[track ? 1 : 0];
  LPC_REQUIRE(per_sm > 0, "kernel does not fit on an SM");
  // enough blocks to fill the chip, no more than there are thread-loads of work
  const long long units = t->dev.n_pad / var.rpt;
  int grid = t->sm_count * per_sm;
  long long want = std::max<long long>(1, (units + TPB - 1) / TPB);
  if(want < grid) grid = (int)want;
  SegTable seg;
  seg.nseg = t->seg_n;
  for(int i = 0; i <= seg.nseg; ++i) seg.q[i] = t->seg_q[i];
  WlState w{};
  unsigned switch_at = 0;
  LPC_CUDA(cudaEventRecord(s->ev0, st));   // device_ms covers the scratch clears as well as the kernel
  if(track) {
    int rc = get_scratch(s, t, &w);
    if(rc) return rc;
    LPC_CUDA(cudaMemsetAsync(w.stamp, 0, std::max<long long>(t->dev.n, 1) * 4, st));
    LPC_CUDA(cudaMemsetAsync(w.vmark, 0, (size_t)std::max(1, s->nvars) * 4, st));
    if(o->mode == LPC_MODE_WORKLIST) switch_at = 0xffffffffu;
    else {
      int div = o->reserved > 0 ? o->reserved : 128;
      switch_at = (unsigned)std::max<long long>(1, t->dev.n / div);
    }
  }
  LPC_CUDA(cudaMemsetAsync(s->d_ctl, 0, sizeof(FixCtl), st));
  TableDev td = t->dev;
  int2* store = s->d;
  FixCtl* ctl = s->d_ctl;
  int max_sweeps = o->max_sweeps, stop = o->stop_on_bot;
  int use_vote = 1;   // LPC_VOTE=0: flag words + cooperative_groups grid.sync() (kept for A/B runs, profiles/r01_summary.md)
  if(const char* e = getenv("LPC_VOTE")) use_vote = atoi(e);
  // a table with at most a couple of units per thread is swept as ONE segment: with a handful of records per block the
  // per-segment passes only add dependent memory round trips (config 1: 3 segments x 2 round trips per sweep)
  if(units <= 2LL * grid * TPB) { seg.nseg = 1; seg.q[0] = 0; seg.q[1] = (int)(t->dev.n_pad / 4); }
  int sm_order = 2;   // LPC_SMORDER=0: fractions in blockIdx order, 1: in SM order, 2: interleaved per SM (A/B runs)
  if(const char* e = getenv("LPC_SMORDER")) sm_order = atoi(e);
  void* args[] = {&td, &store, &seg, &ctl, &w, &max_sweeps, &stop, &switch_at, &use_vote, &sm_order};
  LPC_CUDA(cudaLaunchCooperativeKernel((void*)k, dim3(grid), dim3(TPB), args, 0, st));
  g_launches++;
  LPC_CUDA(cudaEventRecord(s->ev1, st));
  LPC_CUDA(cudaMemcpyAsync(s->h_ctl, s->d_ctl, sizeof(FixCtl), cudaMemcpyDeviceToHost, st));
  s->last_stream = st;
  s->pending = true;
  return LPC_OK;
}

int lpc_fixpoint_collect(lpc_store* s, lpc_fixpoint_result* r) {
  LPC_REQUIRE(s != nullptr, "null store");
  LPC_REQUIRE(s->pending, "no fixpoint in flight on this store");
  LPC_CUDA(cudaStreamSynchronize(s->last_stream));
  s->pending = false;
  if(r) {
    memset(r, 0, sizeof(*r));
    r->has_changed = s->h_ctl->has_changed;
    r->is_bot = s->h_ctl->is_bot;
    r->sweeps = s->h_ctl->sweeps;
    r->dense_sweeps = s->h_ctl->dense_sweeps;
    r->deductions = (int64_t)s->h_ctl->deductions;
    float ms = 0;
    LPC_CUDA(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    r->device_ms = ms;
  }
  return LPC_OK;
}

int lpc_fixpoint(const lpc_table* t, lpc_store* s, const lpc_fixpoint_opts* o, lpc_fixpoint_result* r) {
  int rc = lpc_fixpoint_async(t, s, o);
  if(rc) return rc;
  return lpc_fixpoint_collect(s, r);
}

int lpc_fixpoint_host(const lpc_table* t, int32_t* lbub, const lpc_fixpoint_opts* o, lpc_fixpoint_result* r) {
  LPC_REQUIRE(t && lbub, "null argument");
  // one cached device store per table for the host-buffer entry point
  lpc_table* tm = const_cast<lpc_table*>(t);
  if(!tm->host_store) {
    int rc = lpc_store_create(t->dev.nvars, &tm->host_store);
    if(rc) return rc;
  }
  lpc_store* s = tm->host_store;
  cudaStream_t st = o ? (cudaStream_t)o->stream : nullptr;
  size_t bytes = (size_t)t->dev.nvars * 8;
  if(bytes) LPC_CUDA(cudaMemcpyAsync(s->d, lbub, bytes, cudaMemcpyHostToDevice, st));
  int rc = lpc_fixpoint_async(t, s, o);
  if(rc) return rc;
  if(bytes) LPC_CUDA(cudaMemcpyAsync(lbub, s->d, bytes, cudaMemcpyDeviceToHost, st));
  return lpc_fixpoint_collect(s, r);
}

int lpc_deduce_one(const lpc_table* t, lpc_store* s, int64_t i, int* changed) {
  LPC_REQUIRE(t && s, "null argument");
  LPC_REQUIRE(i >= 0 && i < t->dev.n, "record index out of range");   // assert at pir.hpp:388
  LPC_REQUIRE(s->nvars >= t->dev.nvars, "store smaller than the table's variable range");
  k_deduce_one<<<1, 1>>>(t->dev, s->d, i, &s->d_ctl->scratch[0]);
  g_launches++;
  LPC_CUDA(cudaGetLastError());
  int c = 0;
  LPC_CUDA(cudaMemcpy(&c, &s->d_ctl->scratch[0], sizeof(int), cudaMemcpyDeviceToHost));
  if(changed) *changed = c;
  return LPC_OK;
}

static int ask_impl(const lpc_table* t, const lpc_store* s, int64_t* n_entailed, uint8_t* host_bits) {
  LPC_REQUIRE(t && s, "null argument");
  LPC_REQUIRE(s->nvars >= t->dev.nvars, "store smaller than the table's variable range");
  unsigned long long* d_cnt = nullptr;
  uint8_t* d_bits = nullptr;
  LPC_CUDA(cudaMalloc((void**)&d_cnt, 8));
  LPC_CUDA(cudaMemset(d_cnt, 0, 8));
  if(host_bits && t->dev.n) LPC_CUDA(cudaMalloc((void**)&d_bits, t->dev.n));
  if(t->dev.n) {
    int blocks = (int)std::min<long long>(ceil_div(t->dev.n, 256), 148 * 8);
    k_ask_all<<<blocks, 256>>>(t->dev, s->d, d_cnt, d_bits);
    g_launches++;
    LPC_CUDA(cudaGetLastError());
  }
  unsigned long long c = 0;
  LPC_CUDA(cudaMemcpy(&c, d_cnt, 8, cudaMemcpyDeviceToHost));
  if(host_bits && t->dev.n) LPC_CUDA(cudaMemcpy(host_bits, d_bits, t->dev.n, cudaMemcpyDeviceToHost));
  cudaFree(d_cnt);
  cudaFree(d_bits);
  if(n_entailed) *n_entailed = (int64_t)c;
  return LPC_OK;
}

int lpc_ask_all(const lpc_table* t, const lpc_store* s, int64_t* n_entailed) { return ask_impl(t, s, n_entailed, nullptr); }

int lpc_ask_bits(const lpc_table* t, const lpc_store* s, uint8_t* out) {
  LPC_REQUIRE(out != nullptr || (t && t->dev.n == 0), "null output");
  return ask_impl(t, s, nullptr, out);
}

int lpc_ask_one(const lpc_table* t, const lpc_store* s, int64_t i, int* entailed) {
  LPC_REQUIRE(t && s && entailed, "null argument");
  LPC_REQUIRE(i >= 0 && i < t->dev.n, "record index out of range");
  std::vector<uint8_t> bits(t->dev.n);
  int rc = ask_impl(t, s, nullptr, bits.data());
  if(rc) return rc;
  *entailed = bits[i];
  return LPC_OK;
}

} // extern "C"
