// pir_fixpoint.cu — the single-store fixpoint, dense mode: a persistent cooperative kernel for sm_100a.
//
// Replaces the loop  GaussSeidelIteration{}.fixpoint(pir.num_deductions(), [&](size_t i){ return pir.deduce(i); })
// (call sites tests/pir_test.cpp:60-62, 82-86) with chaotic parallel iteration. Because every propagator is
// monotone, the least fixpoint does not depend on the schedule; only sweep and deduction counts do.
//
// One launch = one fixpoint:
//   dense sweeps   the table is cut into its opcode segments (it is sorted by (op, y, x, z), pir.hpp:343-347) and
//                  every resident block owns the same contiguous fraction of EVERY segment: contiguous so that the
//                  sort order turns into L1 locality of the gathers, per segment so that all blocks see the same
//                  operator mix and reach the grid barrier together. A thread fetches RPT records with vector loads,
//                  gathers their intervals, evaluates the rules in registers and joins tightened bounds with
//                  atomicMax / atomicMin (RED at L2).
//   termination    a vote-carrying grid barrier per sweep (grid_barrier.cuh); its fence drops stale L1 lines, so the
//                  last (quiescent) sweep reads the final store.
// The change-driven modes (LPC_MODE_AUTO, LPC_MODE_WORKLIST) are pir_dirty.cu.
#include "lpc_internal.cuh"
#include "grid_barrier.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace lpc {

constexpr int TPB = 256;
constexpr int MAX_SEG = 16;

// Opcode segments of the padded table, in quads (4 records). nseg == 0 never happens (an empty table has one
// all-padding segment).
struct SegTable {
  int nseg;
  int q[MAX_SEG + 1];
};

// Join the new domain into the store. Returns bit0 = changed, bit1 = became empty, bit2 = overflow hazard (a finite
// bound within 2^24 of the int32 limits was written: from there on `+` and `*` may wrap, lpc.h).
__device__ __forceinline__ int commit(int2* p, int2 old, const Itv& nw) {
  int f = 0;
  if(nw.lb > old.x) { atomicMax(&p->x, nw.lb); f = 1 | (near_inf_lo(nw.lb) ? 4 : 0); }
  if(nw.ub < old.y) { atomicMin(&p->y, nw.ub); f |= 1 | (near_inf_hi(nw.ub) ? 4 : 0); }
  if(f && nw.lb > nw.ub) f |= 2;
  return f;
}

// One propagator evaluation: rules in registers, then (rarely) the joins. The common case - nothing tightens, no
// operand empty - is nine compares folded into one predicate and a single not-taken branch.
template <bool HAS_DIV>
__device__ __forceinline__ int run_record(int op, int xi, int yi, int zi, int2 a, int2 b, int2 c, int2* store) {
  Itv r1(a.x, a.y), r2(b.x, b.y), r3(c.x, c.y);
  deduce_regs<HAS_DIV>(op, r1, r2, r3);
  const bool slow = (r1.lb > a.x) | (r1.ub < a.y) | (r2.lb > b.x) | (r2.ub < b.y) | (r3.lb > c.x) | (r3.ub < c.y)
                  | (a.x > a.y) | (b.x > b.y) | (c.x > c.y);
  int f = 0;
  if(slow) {
    f = ((a.x > a.y) | (b.x > b.y) | (c.x > c.y)) ? 2 : 0;
    f |= commit(store + xi, a, r1);
    f |= commit(store + yi, b, r2);
    f |= commit(store + zi, c, r3);
  }
  return f;
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

template <bool HAS_DIV, int RPT, int MINB>
__global__ void __launch_bounds__(TPB, MINB) k_pir_fixpoint(TableDev t, int2* store, SegTable seg, FixCtl* ctl,
                                                            int max_sweeps, int stop_on_bot, int sm_order) {
  __shared__ unsigned long long s_vote;
  const int tid = threadIdx.x;
  const long long gtid = blockIdx.x * (long long)TPB + tid;
  const long long gthreads = (long long)gridDim.x * TPB;
  int nbar = 0;

  // ---- prologue: a store created with an empty variable is at bot before the first sweep ----
  const int sm_slot = sm_order ? sm_rank_arrive(ctl->sm_slots) : 0;
  int hazard = 0;
  bool bot;
  {
    int f = 0;
    for(long long i = gtid; i < t.nvars; i += gthreads) {
      const int2 v = store[i];
      f |= v.x > v.y;
      hazard |= near_inf_lo(v.x) | near_inf_hi(v.y);
    }
    bot = grid_vote_barrier(ctl->bar, nbar++, false, f != 0, &s_vote).bot;
  }
  // which fraction of every segment this block sweeps: its rank in SM order (co-resident blocks get adjacent fractions)
  // sm_order 1: this block's own fraction; 2: the SM's blocks interleave over the SM's fraction (one moving window)
  long long bid = blockIdx.x, bspan = 1;
  int bslot = 0;
  if(sm_order) {
    const SmRank r = sm_rank_resolve(ctl->sm_slots, sm_slot);
    if(sm_order == 2) { bid = r.below; bspan = r.here; bslot = r.slot; }
    else bid = r.below + r.slot;
  }
  int sweeps = 0;
  bool any_changed = false;
  unsigned long long deductions = 0;
  bool done = (bot && stop_on_bot) || t.n == 0;
  constexpr int UPQ = 4 / RPT;   // units per quad

  while(!done) {
    int f = 0;
    for(int s = 0; s < seg.nseg; ++s) {
      const long long sq0 = seg.q[s], len = seg.q[s + 1] - sq0;
      const int u0 = (int)((sq0 + len * bid / gridDim.x) * UPQ);
      const int u1 = (int)((sq0 + len * (bid + bspan) / gridDim.x) * UPQ);
      for(int u = u0 + bslot * TPB + tid; u < u1; u += (int)bspan * TPB) {
        Unit<RPT> r;
        r.load(t, u);
        int2 a[RPT], b[RPT], c[RPT];
#pragma unroll
        for(int k = 0; k < RPT; ++k) { a[k] = store[r.x[k]]; b[k] = store[r.y[k]]; c[k] = store[r.z[k]]; }
#pragma unroll
        for(int k = 0; k < RPT; ++k) f |= run_record<HAS_DIV>(r.op[k], r.x[k], r.y[k], r.z[k], a[k], b[k], c[k], store);
      }
    }
    hazard |= f & 4;
    // the arrival at the grid barrier carries the block's votes (grid_barrier.cuh): one L2 round trip instead of four
    const GridVote v = grid_vote_barrier(ctl->bar, nbar++, f & 1, f & 2, &s_vote);
    ++sweeps;
    deductions += (unsigned long long)t.n;
    bot |= v.bot;
    any_changed |= v.changed;
    if(!v.changed || (bot && stop_on_bot) || (max_sweeps && sweeps >= max_sweeps)) done = true;
  }
  if(__syncthreads_or(hazard) && tid == 0) atomicOr(&ctl->hazard, 1);
  if(blockIdx.x == 0 && tid == 0) {
    ctl->sweeps = sweeps;
    ctl->dense_sweeps = sweeps;
    ctl->has_changed = any_changed;
    ctl->is_bot = bot;
    ctl->deductions = deductions;
  }
}

// ---- single-step and entailment kernels ----------------------------------------------------------------------------
__global__ void k_deduce_one(TableDev t, int2* store, long long i, int* out) {
  const int op = t.op[i], xi = t.x[i], yi = t.y[i], zi = t.z[i];
  int2 a = store[xi], b = store[yi], c = store[zi];
  int f = run_record<true>(op, xi, yi, zi, a, b, c, store);
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

typedef void (*fix_kernel_t)(TableDev, int2*, SegTable, FixCtl*, int, int, int);

// The (records per thread, min blocks per SM) variants that are built; LPC_RPT / LPC_MINB select one for tuning.
struct Variant { int rpt, minb; fix_kernel_t k[2]; };
#define LPC_VARIANT(R, M) { R, M, { k_pir_fixpoint<false, R, M>, k_pir_fixpoint<true, R, M> } }
static const Variant kVariants[] = { LPC_VARIANT(2, 3), LPC_VARIANT(2, 4), LPC_VARIANT(4, 3), LPC_VARIANT(1, 4) };
static const int kDefaultVariant = 0;   // RPT 2, 3 blocks / SM: fastest on config 2 (profiles/r01_ab_variants.txt)

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

// Handles live on the device they were created on; a call with another current device would launch on foreign memory.
int lpc_check_device(int handle_device, const char* what) {
  int dev = -1;
  LPC_CUDA(cudaGetDevice(&dev));
  if(dev != handle_device) {
    set_error("%s: the handle lives on CUDA device %d, the current device is %d", what, handle_device, dev);
    return LPC_ERR_INVALID;
  }
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
  LPC_REQUIRE(!s->pending, "a fixpoint is still in flight on this store (collect it first)");
  LPC_REQUIRE(t->device == s->device, "table and store live on different CUDA devices");
  int rc = lpc_check_device(s->device, "lpc_fixpoint");
  if(rc) return rc;
  lpc_fixpoint_opts def;
  if(!o) { lpc_fixpoint_default_opts(&def); o = &def; }
  LPC_REQUIRE(o->mode >= LPC_MODE_AUTO && o->mode <= LPC_MODE_WORKLIST, "bad mode");
  cudaStream_t st = (cudaStream_t)o->stream;
  const Variant& var = pick_variant();
  fix_kernel_t k = var.k[t->has_div ? 1 : 0];
  // launch plan: computed once per table (the segment scan is O(n) on the host, the occupancy query is a driver call)
  if(!t->plan_ready) {
    SegTable sg = build_segments(t);
    t->seg_n = sg.nseg;
    for(int i = 0; i <= sg.nseg; ++i) t->seg_q[i] = sg.q[i];
    const char* ce = getenv("LPC_CARVE");   // see pir_dirty.cu
    if(!ce || atoi(ce))
      LPC_CUDA(cudaFuncSetAttribute((const void*)k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1));
    LPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&t->blocks_per_sm[0], k, TPB, 0));
    t->plan_ready = true;
  }
  // The change-driven kernel of pir_dirty.cu: dense sweeps that learn to skip 64-record groups whose variables did not
  // change and strike groups whose propagators are all entailed. LPC_MODE_WORKLIST hands over to flagged sweeps right
  // after the first sweep; LPC_MODE_AUTO once a sweep changes few groups, and leaves a table with at most a couple of
  // units per thread to the plain dense kernel below (such a table is swept in a few microseconds and only pays the
  // bookkeeping: config 1 115 vs 107 us).
  const bool small_table = t->dev.n_pad / 2 <= 2LL * t->sm_count * 3 * TPB;
  if(o->mode == LPC_MODE_WORKLIST || (o->mode == LPC_MODE_AUTO && !small_table)) return lpc_dirty_fixpoint_launch(t, s, o);
  const int per_sm = t->blocks_per_sm[0];
  LPC_REQUIRE(per_sm > 0, "kernel does not fit on an SM");
  // enough blocks to fill the chip, no more than there are thread-loads of work
  const long long units = t->dev.n_pad / var.rpt;
  int grid = t->sm_count * per_sm;
  // enough blocks for the records - and for the prologue's scan of the store (a handful of records over a million
  // variables would otherwise leave that scan to one block: 400 us)
  long long want = std::max<long long>({1, (units + TPB - 1) / TPB, ((long long)s->nvars + 16 * TPB - 1) / (16 * TPB)});
  if(want < grid) grid = (int)want;
  SegTable seg;
  seg.nseg = t->seg_n;
  for(int i = 0; i <= seg.nseg; ++i) seg.q[i] = t->seg_q[i];
  LPC_CUDA(cudaEventRecord(s->ev0, st));
  LPC_CUDA(cudaMemsetAsync(s->d_ctl, 0, sizeof(FixCtl), st));
  TableDev td = t->dev;
  int2* store = s->d;
  FixCtl* ctl = s->d_ctl;
  int max_sweeps = o->max_sweeps, stop = o->stop_on_bot;
  // a table with at most a couple of units per thread is swept as ONE segment: with a handful of records per block the
  // per-segment passes only add dependent memory round trips (config 1: 3 segments x 2 round trips per sweep)
  if(units <= 2LL * grid * TPB) { seg.nseg = 1; seg.q[0] = 0; seg.q[1] = (int)(t->dev.n_pad / 4); }
  int sm_order = 2;   // LPC_SMORDER=0: fractions in blockIdx order, 1: in SM order, 2: interleaved per SM (A/B runs)
  if(const char* e = getenv("LPC_SMORDER")) sm_order = atoi(e);
  void* args[] = {&td, &store, &seg, &ctl, &max_sweeps, &stop, &sm_order};
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
    r->overflow_hazard = s->h_ctl->hazard;
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
  unsigned long long c = 0;
  auto body = [&]() -> int {
    LPC_CUDA(cudaMalloc((void**)&d_cnt, 8));
    LPC_CUDA(cudaMemset(d_cnt, 0, 8));
    if(host_bits && t->dev.n) LPC_CUDA(cudaMalloc((void**)&d_bits, t->dev.n));
    if(t->dev.n) {
      int blocks = (int)std::min<long long>(ceil_div(t->dev.n, 256), 148 * 8);
      k_ask_all<<<blocks, 256>>>(t->dev, s->d, d_cnt, d_bits);
      g_launches++;
      LPC_CUDA(cudaGetLastError());
    }
    LPC_CUDA(cudaMemcpy(&c, d_cnt, 8, cudaMemcpyDeviceToHost));
    if(host_bits && t->dev.n) LPC_CUDA(cudaMemcpy(host_bits, d_bits, t->dev.n, cudaMemcpyDeviceToHost));
    return LPC_OK;
  };
  const int rc = body();
  cudaFree(d_cnt);   // the scratch goes whether or not a step failed
  cudaFree(d_bits);
  if(rc) return rc;
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
