// pir_dirty.cu — the change-driven single-store fixpoint (LPC_MODE_AUTO, LPC_MODE_WORKLIST): dense sweeps that learn
// to skip (sm_100a).
//
// BASELINE.json's north_star asks for "an optional change-driven worklist so only propagators touching changed
// variables are re-run". A record-granular queue (round 1's LPC_MODE_WORKLIST) paid an atomicExch per enqueued record and
// gave up the streaming table scan: on config 2 it ran 4x fewer deductions and took 6.0 ms against 0.44 ms for dense
// sweeps, and was removed. This kernel keeps the dense sweep's shape - same partition, same coalesced record loads, same
// rules - and makes whole 64-record GROUPS the unit of the worklist (LPC_MODE_WORKLIST = hand over to flagged sweeps
// right after the first sweep, LPC_MODE_AUTO = once few groups change):
//   * while many groups change per sweep (the first sweeps) it is the dense kernel: no bookkeeping at all;
//   * once a sweep changes at most 1/8 of the groups, the next sweep still evaluates everything but every tightening of
//     a variable v also marks the groups of v's incident records (var -> records CSR, built by lpc_table_create) in a
//     byte map for the following sweep;
//   * from then on a warp first reads the flags of its groups (one coalesced byte load per 32 iterations) and only
//     evaluates flagged groups, marking as it tightens. Three byte maps rotate: read, write, being cleared.
//   * entailment-driven elimination in EVERY phase: each evaluation also runs PIR::ask
//     (pir.hpp:417-438) on the bounds it just computed, and a group whose 64 records are all entailed is struck from a
//     fourth byte map ("live") for the rest of the fixpoint - the use the reference makes of ask() in
//     deinterpret(env, remove_entailed) (pir.hpp:912-925). An entailed propagator holds on every point of its box, a
//     sound propagator can then remove nothing from any sub-box, and entailment survives tightening, so a struck group
//     could never change the store again. Config 2: 99.7 % of the `*` and 99.9 % of the `<=` records are entailed after
//     one sweep (both factors fixed / relation decided), i.e. half the table is gone from the third sweep on.
// Invariant: a change of v after the last evaluation of a record on v always flags that record's group for the next
// sweep (also when the evaluation read a stale L1 copy), so at the sweep that changes nothing every record has been
// evaluated on the final value of its variables: the store is the common fixpoint, the same as Gauss-Seidel's
// (DESIGN.md §2). Sweeps end in the count-carrying grid barrier of grid_barrier.cuh.
#include "lpc_internal.cuh"
#include "grid_barrier.cuh"

#include <algorithm>
#include <cstring>

namespace lpc {

constexpr int DTPB = 256;
constexpr int D_MAX_SEG = 16;
struct DSeg { int nseg; int u[D_MAX_SEG + 1]; };   // opcode segments in units of 2 records

template <bool HAS_DIV>
__global__ void __launch_bounds__(DTPB, 3) k_pir_dirty(TableDev t, int2* store, DSeg seg, FixCtl* ctl, unsigned char* dmap,
                                                       int n_groups, int map_stride, unsigned switch_groups, int max_sweeps,
                                                       int stop_on_bot, int sm_order, int strided, int ask_every) {
  __shared__ unsigned long long s_vote;
  __shared__ unsigned s_cnt;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long gtid = blockIdx.x * (long long)DTPB + tid;
  const long long gthreads = (long long)gridDim.x * DTPB;
  int nbar = 0;
  int hazard = 0;
  bool bot;
  const int sm_slot = sm_order ? sm_rank_arrive(ctl->sm_slots) : 0;
  {
    int f = 0;
    for(long long i = gtid; i < t.nvars; i += gthreads) { int2 v = store[i]; f |= v.x > v.y; hazard |= near_inf_lo(v.x) | near_inf_hi(v.y); }
    for(long long i = gtid; i < 3LL * map_stride; i += gthreads) dmap[i] = 0;
    for(long long i = gtid; i < map_stride; i += gthreads) dmap[3LL * map_stride + i] = 1;   // every group live
    bot = grid_vote_barrier(ctl->bar, nbar++, false, f != 0, &s_vote).bot;
  }
  // this block's fraction of every segment: its rank in SM order (grid_barrier.cuh)
  long long bid = blockIdx.x;
  if(sm_order) { const SmRank r = sm_rank_resolve(ctl->sm_slots, sm_slot); bid = r.below + r.slot; }
  int sweeps = 0, dense = 0;
  bool any_changed = false;
  unsigned evals_total = 0;             // records this lane evaluated over the whole fixpoint
  unsigned char* live = dmap + 3 * (size_t)map_stride;
  bool done = (bot && stop_on_bot) || t.n == 0;
  int phase = 0;   // 0: dense, no marking; 1: dense + marking; 2: flagged groups only + marking
  while(!done) {
    const unsigned char* dcur = dmap + (size_t)(sweeps % 3) * map_stride;
    unsigned char* dnext = phase ? dmap + (size_t)((sweeps + 1) % 3) * map_stride : nullptr;
    if(phase) {   // the map of the sweep after next: nobody reads or writes it during this sweep
      unsigned char* dclr = dmap + (size_t)((sweeps + 2) % 3) * map_stride;
      for(long long i = gtid * 16; i < map_stride; i += gthreads * 16) *reinterpret_cast<uint4*>(dclr + i) = make_uint4(0, 0, 0, 0);
    }
    if(tid == 0) s_cnt = 0;
    __syncthreads();
    int f = 0;
    unsigned my_groups = 0, my_evals = 0;
    // One warp-step: this lane's unit u (2 records) of a 64-record group, evaluated iff `active`; then the group's flags.
    // entailment is looked for in the first sweeps, where it is found (config 2: the `*` and `<=` records decided by the
    // first sweep), and every fourth sweep afterwards: striking a group is an optimisation, never a need
    const bool do_ask = ask_every <= 1 || sweeps < 3 || (sweeps % ask_every) == 0;
    auto eval_unit = [&](int u, bool active) {
      int g1 = 0;
      bool ent = active && do_ask;                      // both records of this lane entailed on the bounds just computed
      int cv[6] = {-1, -1, -1, -1, -1, -1};   // variables this lane tightened (2 records x 3 operands)
      if(active) {
        const uchar2 o = reinterpret_cast<const uchar2*>(t.op)[u];
        const int2 X = reinterpret_cast<const int2*>(t.x)[u], Y = reinterpret_cast<const int2*>(t.y)[u],
                   Z = reinterpret_cast<const int2*>(t.z)[u];
        const int2 a0v = store[X.x], b0v = store[Y.x], c0v = store[Z.x];
        const int2 a1v = store[X.y], b1v = store[Y.y], c1v = store[Z.y];
#pragma unroll
        for(int h = 0; h < 2; ++h) {
          const int op = h ? o.y : o.x, xi = h ? X.y : X.x, yi = h ? Y.y : Y.x, zi = h ? Z.y : Z.x;
          const int2 a = h ? a1v : a0v, b = h ? b1v : b0v, c = h ? c1v : c0v;
          Itv r1(a.x, a.y), r2(b.x, b.y), r3(c.x, c.y);
          deduce_regs<HAS_DIV>(op, r1, r2, r3);
          if(do_ask) ent &= ask_regs(op, r1, r2, r3);
          const bool slow = (r1.lb > a.x) | (r1.ub < a.y) | (r2.lb > b.x) | (r2.ub < b.y) | (r3.lb > c.x) | (r3.ub < c.y)
                          | (a.x > a.y) | (b.x > b.y) | (c.x > c.y);
          if(slow) {
            if((a.x > a.y) | (b.x > b.y) | (c.x > c.y)) g1 |= 2;
#pragma unroll
            for(int w = 0; w < 3; ++w) {
              const int v = w == 0 ? xi : w == 1 ? yi : zi;
              const int2 old = w == 0 ? a : w == 1 ? b : c;
              const Itv nw = w == 0 ? r1 : w == 1 ? r2 : r3;
              int ch = 0;
              if(nw.lb > old.x) { atomicMax(&store[v].x, nw.lb); ch = 1; hazard |= near_inf_lo(nw.lb); }
              if(nw.ub < old.y) { atomicMin(&store[v].y, nw.ub); ch = 1; hazard |= near_inf_hi(nw.ub); }
              if(ch) {
                g1 |= nw.lb > nw.ub ? 3 : 1;
                cv[h * 3 + w] = v;
              }
            }
          }
        }
      }
      f |= g1;
      const bool grp_changed = __any_sync(0xffffffffu, g1 & 1);
      if(grp_changed && dnext) {
        // flag the groups of the records incident to every tightened variable (var -> records CSR)
        {
          // the warp walks each variable's row together (one coalesced load of up to 32 record indices). Letting every
          // lane walk the rows of its own variables instead was measured slower (0.41-0.43 vs 0.395 ms on config 2,
          // profiles/r01_ab_strided.txt).
#pragma unroll
          for(int q = 0; q < 6; ++q) {
            unsigned m = __ballot_sync(0xffffffffu, cv[q] >= 0);
            if(!m) continue;
            int rb = 0, re = 0;
            if(cv[q] >= 0) { rb = t.inc_off[cv[q]]; re = t.inc_off[cv[q] + 1]; }
            while(m) {
              const int src = __ffs(m) - 1;
              m &= m - 1;
              const int bb = __shfl_sync(0xffffffffu, rb, src), ee = __shfl_sync(0xffffffffu, re, src);
              for(int j = bb + lane; j < ee; j += 32) dnext[t.inc_idx[j] >> 6] = 1;
            }
          }
        }
      }
      if(grp_changed) ++my_groups;
      if(active) my_evals += 2;
      // all 64 records of a fully covered group entailed: strike it (a group cut by a segment or share boundary stays)
      if(__all_sync(0xffffffffu, ent) && lane == 0) live[u >> 5] = 0;
    };
    if(phase == 2 && strided) {
      // Flagged sweeps, groups dealt to the warps of the whole grid round-robin: flagged groups cluster (a change moves
      // along the sort order), and with contiguous shares a few blocks did all the work of a sparse sweep while the rest
      // waited at the barrier. A warp reads the flags of its next 32 groups with one strided byte load + ballot.
      const long long gwarp = gtid >> 5, gwarps = gthreads >> 5;
      const int n_units = (int)(t.n_pad / 2);
      for(long long base = gwarp; base < n_groups; base += 32 * gwarps) {
        const long long g = base + (long long)lane * gwarps;
        unsigned need = __ballot_sync(0xffffffffu, g < n_groups && dcur[g] != 0 && live[g] != 0);
        while(need) {
          const int k = __ffs(need) - 1;
          need &= need - 1;
          const int u = (int)(base + (long long)k * gwarps) * 32 + lane;
          eval_unit(u, u < n_units);
        }
      }
    }
    else
    for(int s = 0; s < seg.nseg; ++s) {
      // this block's contiguous share of the segment, cut at group boundaries (32 units = 64 records)
      const long long s0 = seg.u[s], s1 = seg.u[s + 1], len = s1 - s0;
      long long b0 = s0 + len * bid / gridDim.x, b1 = s0 + len * (bid + 1) / gridDim.x;
      b0 = bid == 0 ? s0 : max(s0, b0 & ~31LL);
      b1 = bid == gridDim.x - 1 ? s1 : max(s0, b1 & ~31LL);
      const int u0 = (int)b0, u1 = (int)b1;
      if(u0 >= u1) continue;
      const int a0 = u0 & ~31;
      const int niter = (u1 - a0 + DTPB - 1) / DTPB;
      for(int kb = 0; kb < niter; kb += 32) {
        unsigned need;
        {
          const int k = kb + lane;
          const int g = (a0 + warp * 32 + k * DTPB) >> 5;
          const bool in = k < niter && g < n_groups;
          const int fl = in ? (live[g] != 0 && (phase < 2 || dcur[g] != 0)) : 0;
          need = __ballot_sync(0xffffffffu, fl != 0);
        }
        const int kend = min(niter, kb + 32);
        for(int k = kb; k < kend; ++k) {
          if(!((need >> (k - kb)) & 1u)) continue;
          const int u = a0 + tid + k * DTPB;
          eval_unit(u, u >= u0 && u < u1);
        }
      }
    }
    // changed groups of the block -> the barrier's count; evaluated groups -> the deduction counter
    if(lane == 0 && my_groups) atomicAdd(&s_cnt, my_groups);
    const bool any_bot = __syncthreads_or(f & 2);
    const GridVote v = grid_count_barrier(ctl->bar, nbar++, s_cnt, any_bot, &s_vote);
    ++sweeps;
    if(phase < 2) ++dense;
    evals_total += my_evals;
    bot |= v.bot;
    any_changed |= v.changed;
    if(!v.changed || (bot && stop_on_bot) || (max_sweeps && sweeps >= max_sweeps)) done = true;
    else if(phase == 0) { if(v.n_changed <= switch_groups) phase = 1; }
    else phase = 2;
  }
  // deduce() evaluations executed (one atomic per warp for the whole fixpoint)
  if(__syncthreads_or(hazard) && tid == 0) atomicOr(&ctl->hazard, 1);
  evals_total = __reduce_add_sync(0xffffffffu, evals_total);
  if(lane == 0 && evals_total) atomicAdd(&ctl->deductions, (unsigned long long)evals_total);
  if(blockIdx.x == 0 && tid == 0) {
    ctl->sweeps = sweeps;
    ctl->dense_sweeps = dense;
    ctl->has_changed = any_changed;
    ctl->is_bot = bot;
  }
}

} // namespace lpc

using namespace lpc;

// Called by lpc_fixpoint_async for LPC_MODE_AUTO. The three byte maps live with the store handle.
int lpc_dirty_fixpoint_launch(lpc_table* t, lpc_store* s, const lpc_fixpoint_opts* o) {
  cudaStream_t st = (cudaStream_t)o->stream;
  int rc = lpc_table_ensure_csr(t);   // the var -> records index is built when a change-driven kernel first needs it
  if(rc) return rc;
  const long long n_pad = t->dev.n_pad;
  const int n_groups = (int)((n_pad + 63) / 64);
  const int map_stride = (n_groups + 15) / 16 * 16;
  if(s->dirty_cap < 4LL * map_stride) {   // three rotating change maps + the live map
    cudaFree(s->d_dirty);
    s->d_dirty = nullptr; s->dirty_cap = 0;
    LPC_CUDA(cudaMalloc((void**)&s->d_dirty, 4 * (size_t)map_stride));
    s->dirty_cap = 4LL * map_stride;
  }
  if(!t->dirty_ready) {
    // the kernel's working set is the store window of its SM's table fractions: all of the unified L1 / shared memory
    // array as L1 (LPC_CARVE=0 leaves the driver's default)
    const char* ce = getenv("LPC_CARVE");
    if(!ce || atoi(ce)) {
      LPC_CUDA(cudaFuncSetAttribute(k_pir_dirty<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1));
      LPC_CUDA(cudaFuncSetAttribute(k_pir_dirty<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1));
    }
    for(int d = 0; d < 2; ++d)
      LPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&t->dirty_blocks_per_sm[d], d ? k_pir_dirty<true> : k_pir_dirty<false>, DTPB, 0));
    t->dirty_ready = true;
  }
  const int per_sm = t->dirty_blocks_per_sm[t->has_div ? 1 : 0];
  LPC_REQUIRE(per_sm > 0, "kernel does not fit on an SM");
  int grid = t->sm_count * per_sm;
  const long long units = n_pad / 2;
  grid = (int)std::max<long long>(1, std::min<long long>(grid, std::max<long long>((units + DTPB - 1) / DTPB, ((long long)s->nvars + 16 * DTPB - 1) / (16 * DTPB))));
  DSeg seg;
  seg.nseg = t->seg_n;
  for(int i = 0; i <= t->seg_n; ++i) seg.u[i] = t->seg_q[i] * 2;
  if(units <= 2LL * grid * DTPB) { seg.nseg = 1; seg.u[0] = 0; seg.u[1] = (int)units; }   // small table: one pass (see pir_fixpoint.cu)
  // hand over to flagged sweeps once a sweep changes at most 1/d of the groups (opts.reserved = d, default 8);
  // LPC_MODE_WORKLIST: after the first sweep whatever it changed
  const int div = o->reserved > 0 ? o->reserved : 8;
  unsigned switch_groups = o->mode == LPC_MODE_WORKLIST ? 0xffffffffu : (unsigned)std::max(1, n_groups / div);
  LPC_CUDA(cudaEventRecord(s->ev0, st));
  LPC_CUDA(cudaMemsetAsync(s->d_ctl, 0, sizeof(FixCtl), st));
  TableDev td = t->dev;
  int2* store = s->d;
  FixCtl* ctl = s->d_ctl;
  unsigned char* dmap = s->d_dirty;
  int ng = n_groups, ms = map_stride, max_sweeps = o->max_sweeps, stop = o->stop_on_bot;
  int sm_order = 1;   // LPC_SMORDER=0: fractions in blockIdx order (A/B runs)
  if(const char* e = getenv("LPC_SMORDER")) sm_order = atoi(e);
  int strided = 1;    // LPC_DIRTY_STRIDED=0: flagged sweeps keep the contiguous block shares (A/B runs)
  if(const char* e = getenv("LPC_DIRTY_STRIDED")) strided = atoi(e);
  int ask_every = 4;
  if(const char* e = getenv("LPC_ASK_EVERY")) ask_every = atoi(e);
  void* args[] = {&td, &store, &seg, &ctl, &dmap, &ng, &ms, &switch_groups, &max_sweeps, &stop, &sm_order, &strided, &ask_every};
  void* k = t->has_div ? (void*)k_pir_dirty<true> : (void*)k_pir_dirty<false>;
  LPC_CUDA(cudaLaunchCooperativeKernel(k, dim3(grid), dim3(DTPB), args, 0, st));
  g_launches++;
  LPC_CUDA(cudaEventRecord(s->ev1, st));
  LPC_CUDA(cudaMemcpyAsync(s->h_ctl, s->d_ctl, sizeof(FixCtl), cudaMemcpyDeviceToHost, st));
  s->last_stream = st;
  s->pending = true;
  return LPC_OK;
}
