// pir_batch.cu — batched mode: one store copy per subproblem, one thread block per store (sm_100a).
//
// The design the reference prepares for with `deps.is_shared_copy()` (pir.hpp:182-195: all blocks share one
// read-only bytecode table) taken to its B200 conclusion:
//   * persistent blocks, one per SM when the table is staged in shared memory; the table (13 B per record SoA) is
//     brought in ONCE per block by TMA bulk copies (cp.async.bulk + mbarrier) and stays resident while the block
//     works through hundreds of stores;
//   * stores stream through a 2-deep shared-memory ring: while store k is iterated to its fixpoint, store k+1 is
//     already in flight (cp.async.bulk global->shared) and store k-1 is being written back (shared->global bulk
//     group), so the copy engine, not the threads, moves the 16 KB images;
//   * the fixpoint itself never leaves the SM: bounds are read from shared memory, tightened with shared-memory
//     atomicMax / atomicMin (only when they tighten), has_changed / bot are block votes (__syncthreads_or);
//   * per store: bot flag, all-entailed flag (the ask loop of is_extractable, pir.hpp:873-884), lb(objective);
//     per batch: 4 x int64 reduction record, the payload of the single NCCL all-reduce of the multi-GPU driver.
#include "lpc_internal.cuh"
#include "smem_tma.cuh"
#include "batch_internal.cuh"

#include <algorithm>
#include <cstring>
#include <cstdlib>

namespace lpc {



// ---- one dense sweep of the table over the store at shared address a_S ------------------------------------------------
// The records of one opcode run, with the opcode a compile-time constant: no per-record opcode load, no dispatch, and the
// rule is the only code in the loop body (a sorted table has one run per operator, so warps are uniform anyway; what this
// removes is the ~16 instructions per record of fetching the opcode byte, the jump table and the phi moves after it -
// a third of the instructions of an `x = y + z` record in a kernel that is bound by instruction issue).
// TAB: where the table is read from - 0 global memory (int32 indices), 1 shared memory (int32 indices), 2 shared memory
// as 16-bit byte offsets (TableDev::x16; a_x / a_y / a_z then address 2-byte entries)
template <int OP, bool HAS_DIV, int TAB, bool BF>
__device__ __forceinline__ int sweep_run(int s0, int s1, unsigned a_S, const int* sx, const int* sy, const int* sz,
                                         const uint8_t* sop, unsigned a_x, unsigned a_y, unsigned a_z, unsigned a_op,
                                         int tid, int nthr) {
  int f = 0;
  unsigned macc = 0;
  bool bacc = false;
  for(int i = s0 + tid; i < s1; i += nthr) {
    int op = OP;
    unsigned ax, ay, az;
    if(TAB == 2) {
      if(OP < 0) op = lds_u8(a_op + i);
      ax = a_S + lds_u16(a_x + 2 * i); ay = a_S + lds_u16(a_y + 2 * i); az = a_S + lds_u16(a_z + 2 * i);
    }
    else {
      int xi, yi, zi;
      if(TAB == 1) { if(OP < 0) op = lds_u8(a_op + i); xi = lds_s32(a_x + 4 * i); yi = lds_s32(a_y + 4 * i); zi = lds_s32(a_z + 4 * i); }
      else { if(OP < 0) op = sop[i]; xi = sx[i]; yi = sy[i]; zi = sz[i]; }
      ax = a_S + 8u * xi; ay = a_S + 8u * yi; az = a_S + 8u * zi;
    }
    const int2 a = lds_itv(ax), bb = lds_itv(ay), c = lds_itv(az);
    Itv r1(a.x, a.y), r2(bb.x, bb.y), r3(c.x, c.y);
    deduce_regs<HAS_DIV>(op, r1, r2, r3);
    if(!BF) {   // few changes expected (the nodes of a search): one test, then the rare join path
      const bool slow = (r1.lb > a.x) | (r1.ub < a.y) | (r2.lb > bb.x) | (r2.ub < bb.y) | (r3.lb > c.x) | (r3.ub < c.y)
                      | (a.x > a.y) | (bb.x > bb.y) | (c.x > c.y);
      if(slow) {
        if((a.x > a.y) | (bb.x > bb.y) | (c.x > c.y)) f |= 2;
        f |= commit_smem(ax, a, r1) | commit_smem(ay, bb, r2) | commit_smem(az, c, r3);
      }
      continue;
    }
    // Join without a test per bound. On the batched workloads most warp-iterations tighten something (config 4: 79 % of
    // the `+` ones, ncu source counters), so the flags come from bit arithmetic and the join is ONE divergent region per
    // record: a lane whose record moved anything issues all six shared-memory reductions - those of the bounds that did
    // not move are no-ops of the lattice join. (Six individually predicated reductions were tried first: ptxas wraps each
    // in its own BSSY / BRA / BSYNC, a third of the kernel's instructions were control flow; 6.35 -> 6.10 ms.)
    // The rules return the meet with the old domain (lb only grows, ub only shrinks), so "some bound differs" is a
    // change, and "lb > ub afterwards" covers an operand that was already empty as well as one that just became empty.
    const unsigned moved = (unsigned)((r1.lb ^ a.x) | (r1.ub ^ a.y) | (r2.lb ^ bb.x) | (r2.ub ^ bb.y) | (r3.lb ^ c.x) | (r3.ub ^ c.y));
    if(moved) {
      reds_max(ax, r1.lb); reds_min(ax + 4, r1.ub);
      reds_max(ay, r2.lb); reds_min(ay + 4, r2.ub);
      reds_max(az, r3.lb); reds_min(az + 4, r3.ub);
    }
    macc |= moved;                                                       // flags are assembled once, after the loop
    bacc |= (r1.lb > r1.ub) | (r2.lb > r2.ub) | (r3.lb > r3.ub);
  }
  return f | (macc != 0u ? 1 : 0) | (bacc ? 2 : 0);
}
// OP = -1: opcode read per record (the division operators, and tables that are not sorted by opcode)
template <bool HAS_DIV, int TAB, bool BF>
__device__ __forceinline__ int sweep_table(const OpSegs& segs, int npad, unsigned a_S, const int* sx, const int* sy, const int* sz,
                                           const uint8_t* sop, unsigned a_x, unsigned a_y, unsigned a_z, unsigned a_op,
                                           int tid, int nthr) {
  if(segs.n == 0) return sweep_run<-1, HAS_DIV, TAB, BF>(0, npad, a_S, sx, sy, sz, sop, a_x, a_y, a_z, a_op, tid, nthr);
  int f = 0;
  for(int s = 0; s < segs.n; ++s) {
    const int s0 = segs.start[s], s1 = segs.start[s + 1];
#define LPC_RUN(O) case O: f |= sweep_run<O, HAS_DIV, TAB, BF>(s0, s1, a_S, sx, sy, sz, sop, a_x, a_y, a_z, a_op, tid, nthr); break;
    switch(segs.op[s]) {
      LPC_RUN(D_ADD) LPC_RUN(D_MUL) LPC_RUN(D_MIN) LPC_RUN(D_MAX) LPC_RUN(D_EQ) LPC_RUN(D_LEQ)
      default: f |= sweep_run<-1, HAS_DIV, TAB, BF>(s0, s1, a_S, sx, sy, sz, sop, a_x, a_y, a_z, a_op, tid, nthr); break;
    }
#undef LPC_RUN
  }
  return f;
}

// ---- change-driven fixpoint of ONE store by its block ---------------------------------------------------------------
// The loop of k_pir_batch made incremental: records are evaluated in groups of 32 (one warp-iteration) and a group is
// evaluated in a sweep only if it is flagged; tightening a variable flags, for the next sweep, the groups of its incident
// records (var -> records CSR of lpc_table_create). Three byte maps in shared memory rotate: read, write, being cleared.
// The first sweep takes its flags from `seed_mode`: 0 = every group (an arbitrary store), 1 = the groups incident to the
// `nseeds` variables in `seeds` (a store that is a fixpoint of the table except on those variables: the children of a
// search node, an EPS split of a root fixpoint). The termination argument is DESIGN.md §2's: every change to a variable
// after the last evaluation of a record on it flags that record's group, so when a sweep changes nothing every record
// has been evaluated on the final value of its variables.
template <bool HAS_DIV, bool TABLE_SMEM>
__device__ __forceinline__ bool block_fixpoint_cd(const TableDev& t, const int2* S, unsigned a_S, const int* sx, const int* sy,
                                                  const int* sz, const uint8_t* sop, unsigned a_x, unsigned a_y, unsigned a_z,
                                                  unsigned a_op, volatile int* s_bot, int* s_evals, unsigned char* dmap, int ng,
                                                  int ngp, int seed_mode, const int* seeds, int nseeds, int max_sweeps,
                                                  int stop_on_bot, int& sweeps_out) {
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
  const int npad = (int)t.n_pad;
  if(tid == 0) { s_bot[0] = 0; s_bot[4] = 0; *s_evals = 0; }   // two bot words, one per sweep parity (see k_pir_batch)
  for(int i = tid; i < 3 * ngp; i += nthr) dmap[i] = (seed_mode == 0 && i < ng) ? 1 : 0;
  int f0 = 0;
  for(int v = tid; v < t.nvars; v += nthr) { int2 d = S[v]; f0 |= d.x > d.y; }
  bool bot = __syncthreads_or(f0) != 0;
  if(seed_mode == 1) {
    for(int q = tid; q < nseeds; q += nthr) {
      const int v = seeds[q];
      for(int j = __ldg(&t.inc_off[v]), e = __ldg(&t.inc_off[v + 1]); j < e; ++j) dmap[__ldg(&t.inc_idx[j]) >> 5] = 1;
    }
    __syncthreads();
  }
  int sweeps = 0, evals = 0;
  bool changed = !(bot && stop_on_bot) && t.n > 0;
  while(changed) {
    const unsigned char* cur = dmap + (sweeps % 3) * ngp;
    unsigned char* nxt = dmap + ((sweeps + 1) % 3) * ngp;
    unsigned char* clr = dmap + ((sweeps + 2) % 3) * ngp;
    for(int i = tid; i < ngp; i += nthr) clr[i] = 0;
    int f = 0;
    for(int g = warp; g < ng; g += nwarps) {
      if(!cur[g]) continue;
      ++evals;
      const int i = g * 32 + lane;
      if(i >= npad) continue;
      int op, xi, yi, zi;
      if(TABLE_SMEM) { op = lds_u8(a_op + i); xi = lds_s32(a_x + 4 * i); yi = lds_s32(a_y + 4 * i); zi = lds_s32(a_z + 4 * i); }
      else { op = sop[i]; xi = sx[i]; yi = sy[i]; zi = sz[i]; }
      const unsigned ax = a_S + 8u * xi, ay = a_S + 8u * yi, az = a_S + 8u * zi;
      const int2 a = lds_itv(ax), bb = lds_itv(ay), c = lds_itv(az);
      Itv r1(a.x, a.y), r2(bb.x, bb.y), r3(c.x, c.y);
      deduce_regs<HAS_DIV>(op, r1, r2, r3);
      const bool slow = (r1.lb > a.x) | (r1.ub < a.y) | (r2.lb > bb.x) | (r2.ub < bb.y) | (r3.lb > c.x) | (r3.ub < c.y)
                      | (a.x > a.y) | (bb.x > bb.y) | (c.x > c.y);
      if(slow) {
        if((a.x > a.y) | (bb.x > bb.y) | (c.x > c.y)) f |= 2;
#pragma unroll
        for(int w = 0; w < 3; ++w) {
          const int v = w == 0 ? xi : w == 1 ? yi : zi;
          const int g1 = w == 0 ? commit_smem(ax, a, r1) : w == 1 ? commit_smem(ay, bb, r2) : commit_smem(az, c, r3);
          f |= g1;
          if(g1 & 1) for(int j = __ldg(&t.inc_off[v]), e = __ldg(&t.inc_off[v + 1]); j < e; ++j) nxt[__ldg(&t.inc_idx[j]) >> 5] = 1;
        }
      }
    }
    volatile int* sb = s_bot + 4 * (sweeps & 1);
    ++sweeps;
    if(f & 2) *sb = 1;
    const int any_chg = __syncthreads_or(f & 1);
    bot |= *sb != 0;
    changed = any_chg && !(bot && stop_on_bot) && !(max_sweeps && sweeps >= max_sweeps);
  }
  if(lane == 0 && evals) atomicAdd(s_evals, evals);
  __syncthreads();
  sweeps_out = sweeps;
  return bot;
}

// Shared memory carve-up (dynamic): [mbarriers 64 B][store ring 2 x sbytes][table: x | y | z | op][CD: 3 group maps]
template <bool HAS_DIV, bool TABLE_SMEM, bool CD>
__global__ void k_pir_batch(TableDev t, OpSegs segs, int2* stores, int n_stores, int sbytes, uint8_t* flags, int* sweeps_out,
                            int* obj_out, BatchCtl* ctl, int objective_var, int max_sweeps, int stop_on_bot,
                            const int* seeds, int nseeds) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem);   // [0],[1]: store ring, [2]: table
  int* s_next = reinterpret_cast<int*>(smem + 32);
  volatile int* s_bot = reinterpret_cast<volatile int*>(smem + 36);
  int* s_evals = reinterpret_cast<int*>(smem + 44);
  int2* ring[2] = {reinterpret_cast<int2*>(smem + 64), reinterpret_cast<int2*>(smem + 64 + sbytes)};
  const int* sx = t.x; const int* sy = t.y; const int* sz = t.z; const uint8_t* sop = t.op;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int npad = (int)t.n_pad;
  const size_t store_stride = (size_t)t.nvars;   // in int2

  if(tid == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    fence_mbar_init();
  }
  __syncthreads();
  int cur = blockIdx.x < n_stores ? blockIdx.x : -1;
  if(tid == 0) {
    if(TABLE_SMEM) {
      char* tb = reinterpret_cast<char*>(smem + 64 + 2 * (size_t)sbytes);
      mbar_expect_tx(&bars[2], (unsigned)(npad * 13));
      bulk_g2s_chunked(tb, (const char*)t.x, npad * 4, &bars[2]);
      bulk_g2s_chunked(tb + (size_t)npad * 4, (const char*)t.y, npad * 4, &bars[2]);
      bulk_g2s_chunked(tb + (size_t)npad * 8, (const char*)t.z, npad * 4, &bars[2]);
      bulk_g2s_chunked(tb + (size_t)npad * 12, (const char*)t.op, npad, &bars[2]);
    }
    if(cur >= 0) {
      mbar_expect_tx(&bars[0], (unsigned)sbytes);
      bulk_g2s_chunked((char*)ring[0], (const char*)(stores + cur * store_stride), sbytes, &bars[0]);
    }
  }
  if(TABLE_SMEM) {
    char* tb = reinterpret_cast<char*>(smem + 64 + 2 * (size_t)sbytes);
    sx = reinterpret_cast<const int*>(tb);
    sy = reinterpret_cast<const int*>(tb + (size_t)npad * 4);
    sz = reinterpret_cast<const int*>(tb + (size_t)npad * 8);
    sop = reinterpret_cast<const uint8_t*>(tb + (size_t)npad * 12);
    mbar_wait(&bars[2], 0);
  }
  const unsigned a_x = smem_u32(sx), a_y = smem_u32(sy), a_z = smem_u32(sz), a_op = smem_u32(sop);   // valid iff TABLE_SMEM

  // block-level accumulators (thread 0)
  long long a_sol = 0, a_bot = 0, a_unk = 0, a_sweeps = 0, a_ded = 0;
  int a_best = LPC_INF, a_maxsw = 0;
  unsigned phase[2] = {0, 0};
  int b = 0;
  while(cur >= 0) {
    // claim the next store and start fetching it into the other ring slot
    if(tid == 0) {
      int nx = atomicAdd(&ctl->next_store, 1);
      if(nx >= n_stores) nx = -1;
      *s_next = nx;
      if(nx >= 0) {
        bulk_wait_read0();   // the write-back that last used ring[b^1] has finished reading it
        mbar_expect_tx(&bars[b ^ 1], (unsigned)sbytes);
        bulk_g2s_chunked((char*)ring[b ^ 1], (const char*)(stores + nx * store_stride), sbytes, &bars[b ^ 1]);
      }
    }
    mbar_wait(&bars[b], phase[b]);
    phase[b] ^= 1;
    int2* S = ring[b];
    const unsigned a_S = smem_u32(S);

    bool bot;
    int sweeps = 0;
    long long ded_store;
    if constexpr(CD) {
      const int ng = (npad + 31) / 32, ngp = (ng + 15) / 16 * 16;
      unsigned char* dmap = smem + 64 + 2 * (size_t)sbytes + (TABLE_SMEM ? (size_t)npad * 13 : 0);
      bot = block_fixpoint_cd<HAS_DIV, TABLE_SMEM>(t, S, a_S, sx, sy, sz, sop, a_x, a_y, a_z, a_op, s_bot, s_evals, dmap, ng, ngp,
                                                   nseeds >= 0 ? 1 : 0, seeds, nseeds, max_sweeps, stop_on_bot, sweeps);
      ded_store = 32LL * *s_evals;
    }
    else {
      // bot before the first sweep?
      // Two bot words (smem + 36 and + 52), one per sweep parity: a thread that has left the sweep-ending barrier and runs
      // ahead into the next sweep must not change what a slower thread is about to read as the outcome of the sweep
      // just ended - else the two disagree on whether the store has failed and meet different barriers.
      if(tid == 0) { s_bot[0] = 0; s_bot[4] = 0; }
      int f0 = 0;
      for(int v = tid; v < t.nvars; v += nthr) { int2 d = S[v]; f0 |= d.x > d.y; }
      bot = __syncthreads_or(f0) != 0;
      bool changed = !(bot && stop_on_bot) && t.n > 0;
      while(changed) {
        const int f = sweep_table<HAS_DIV, TABLE_SMEM ? 1 : 0, true>(segs, npad, a_S, sx, sy, sz, sop, a_x, a_y, a_z, a_op, tid, nthr);
        volatile int* sb = s_bot + 4 * (sweeps & 1);
        ++sweeps;
        if(f & 2) *sb = 1;
        // one barrier per sweep: it publishes the sweep's shared-memory joins, votes has_changed, and orders the bot word
        const int any_chg = __syncthreads_or(f & 1);
        bot |= *sb != 0;
        changed = any_chg && !(bot && stop_on_bot) && !(max_sweeps && sweeps >= max_sweeps);
      }
      ded_store = (long long)sweeps * t.n;
    }
    // entailment: the ask loop of is_extractable
    int all_ent = 0;
    if(!bot) {
      int ok = 1;
      for(int i = tid; i < npad && ok; i += nthr) {
        const int2 a = S[sx[i]], bb = S[sy[i]], c = S[sz[i]];
        ok = ask_regs(sop[i], Itv(a.x, a.y), Itv(bb.x, bb.y), Itv(c.x, c.y));
      }
      all_ent = __syncthreads_and(ok);
    }
    // write back: generic-proxy writes -> async proxy, then one bulk store
    fence_async_smem();
    __syncthreads();
    if(tid == 0) {
      for(int o = 0; o < sbytes; o += 32768) bulk_s2g((char*)(stores + cur * store_stride) + o, (char*)S + o, min(32768, sbytes - o));
      bulk_commit();
      flags[cur] = (uint8_t)((bot ? 1 : 0) | (all_ent ? 2 : 0));
      sweeps_out[cur] = sweeps;
      int olb = objective_var >= 0 ? S[objective_var].x : LPC_INF;
      if(obj_out) obj_out[cur] = olb;
      if(bot) ++a_bot; else if(all_ent) ++a_sol; else ++a_unk;
      if(!bot && objective_var >= 0) a_best = min(a_best, olb);
      a_sweeps += sweeps;
      a_ded += ded_store;
      a_maxsw = max(a_maxsw, sweeps);
    }
    cur = *s_next;
    __syncthreads();   // everyone has read s_next before thread 0 overwrites it
    b ^= 1;
  }
  if(tid == 0) {
    bulk_wait0();
    if(a_sol) atomicAdd((unsigned long long*)&ctl->red[0], (unsigned long long)a_sol);
    if(a_bot) atomicAdd((unsigned long long*)&ctl->red[1], (unsigned long long)a_bot);
    if(a_unk) atomicAdd((unsigned long long*)&ctl->red[2], (unsigned long long)a_unk);
    atomicMin(&ctl->red[3], (long long)a_best);
    atomicAdd((unsigned long long*)&ctl->sweeps_total, (unsigned long long)a_sweeps);
    atomicAdd((unsigned long long*)&ctl->deductions, (unsigned long long)a_ded);
    atomicMax(&ctl->max_sweeps_seen, a_maxsw);
  }
}

// ---- in-kernel search: propagate + branch on one store per block ------------------------------------------------------
// snapshot / restore (pir.hpp:857-870) and branching moved next to the fixpoint, so that a
// subproblem is SOLVED by its block instead of only propagated once. Depth-first, deterministic: the variable is the
// first non-singleton of `bvars` (FlatZinc input_order), the value split is a bisection, lower half first
// (indomain_split); a leaf is a store whose branching variables are all fixed - a solution iff every propagator is
// entailed (is_extractable, pir.hpp:873-884). The right branch of every decision is a full store image pushed on a
// per-block stack in global memory by one bulk copy (the device counterpart of snapshot()), popped back by another
// (restore()). Node, failure and solution counts do not depend on the schedule because each node's fixpoint does not.
struct SearchCtl {
  long long n_solutions, n_nodes, n_fails, n_unknown_leaves, n_incomplete, sweeps_total, deductions;
  long long best;      // min over solutions of lb(objective)
  int max_depth_seen;
  int next_store;
};

// One fixpoint of the store at shared address a_S (the loop of k_pir_batch). Returns bot; adds to sweeps.
template <bool HAS_DIV, bool TABLE_SMEM>
__device__ __forceinline__ bool block_fixpoint(const TableDev& t, const OpSegs& segs, const int2* S, unsigned a_S, const int* sx, const int* sy,
                                               const int* sz, const uint8_t* sop, unsigned a_x, unsigned a_y, unsigned a_z,
                                               unsigned a_op, volatile int* s_bot, long long& sweeps_acc) {
  const int tid = threadIdx.x, nthr = blockDim.x, npad = (int)t.n_pad;
  if(tid == 0) { s_bot[0] = 0; s_bot[4] = 0; }   // two bot words, one per sweep parity (see k_pir_batch)
  int f0 = 0;
  for(int v = tid; v < t.nvars; v += nthr) { int2 d = S[v]; f0 |= d.x > d.y; }
  bool bot = __syncthreads_or(f0) != 0;
  bool changed = !bot && t.n > 0;
  int sweeps = 0;
  while(changed) {
    const int f = sweep_table<HAS_DIV, TABLE_SMEM ? 1 : 0, false>(segs, npad, a_S, sx, sy, sz, sop, a_x, a_y, a_z, a_op, tid, nthr);
    volatile int* sb = s_bot + 4 * (sweeps & 1);
    ++sweeps;
    if(f & 2) *sb = 1;
    const int any_chg = __syncthreads_or(f & 1);
    bot |= *sb != 0;
    changed = any_chg && !bot;
  }
  sweeps_acc += sweeps;
  return bot;
}

// Shared memory carve-up (dynamic): [mbarriers + scalars 64 B][store sbytes][table: x | y | z | op][CD: 3 group maps]
template <bool HAS_DIV, bool TABLE_SMEM, bool CD>
__global__ void __launch_bounds__(HAS_DIV ? 256 : 1024) k_pir_search(TableDev t, OpSegs segs, const int2* roots, int n_stores, int sbytes, int2* stack, int* vstack, int max_depth,
                             const int* bvars, int nb, int objective_var, long long max_nodes, long long* per_store,
                             SearchCtl* ctl, const int* root_seeds, int n_root_seeds) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem);   // [0]: store, [2]: table
  int* s_next = reinterpret_cast<int*>(smem + 32);
  volatile int* s_bot = reinterpret_cast<volatile int*>(smem + 36);
  int* s_idx = reinterpret_cast<int*>(smem + 40);
  int* s_evals = reinterpret_cast<int*>(smem + 44);
  int* s_seed = reinterpret_cast<int*>(smem + 48);
  int2* S = reinterpret_cast<int2*>(smem + 64);
  const int* sx = t.x; const int* sy = t.y; const int* sz = t.z; const uint8_t* sop = t.op;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int npad = (int)t.n_pad;
  const size_t store_stride = (size_t)t.nvars;
  int2* my_stack = stack + (size_t)blockIdx.x * max_depth * store_stride;
  int* my_vstack = vstack + (size_t)blockIdx.x * max_depth;
  const int ng = (npad + 31) / 32, ngp = (ng + 15) / 16 * 16;
  unsigned char* dmap = smem + 64 + (size_t)sbytes + (TABLE_SMEM ? (size_t)npad * 13 : 0);
  if(tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[2], 1); fence_mbar_init(); }
  __syncthreads();
  if(TABLE_SMEM) {
    char* tb = reinterpret_cast<char*>(smem + 64 + (size_t)sbytes);
    if(tid == 0) {
      mbar_expect_tx(&bars[2], (unsigned)(npad * 13));
      bulk_g2s_chunked(tb, (const char*)t.x, npad * 4, &bars[2]);
      bulk_g2s_chunked(tb + (size_t)npad * 4, (const char*)t.y, npad * 4, &bars[2]);
      bulk_g2s_chunked(tb + (size_t)npad * 8, (const char*)t.z, npad * 4, &bars[2]);
      bulk_g2s_chunked(tb + (size_t)npad * 12, (const char*)t.op, npad, &bars[2]);
    }
    sx = reinterpret_cast<const int*>(tb);
    sy = reinterpret_cast<const int*>(tb + (size_t)npad * 4);
    sz = reinterpret_cast<const int*>(tb + (size_t)npad * 8);
    sop = reinterpret_cast<const uint8_t*>(tb + (size_t)npad * 12);
    mbar_wait(&bars[2], 0);
  }
  const unsigned a_x = smem_u32(sx), a_y = smem_u32(sy), a_z = smem_u32(sz), a_op = smem_u32(sop);
  const unsigned a_S = smem_u32(S);
  unsigned phase = 0;
  long long a_sol = 0, a_nodes = 0, a_fails = 0, a_unk = 0, a_inc = 0, a_sweeps = 0, a_evals = 0, a_best = LPC_INF;
  int a_maxd = 0;
  int cur = blockIdx.x < n_stores ? blockIdx.x : -1;
  while(cur >= 0) {
    // restore(root): the subproblem's store
    fence_async_smem();
    __syncthreads();
    if(tid == 0) {
      mbar_expect_tx(&bars[0], (unsigned)sbytes);
      bulk_g2s_chunked((char*)S, (const char*)(roots + cur * store_stride), sbytes, &bars[0]);
    }
    mbar_wait(&bars[0], phase); phase ^= 1;
    long long sol = 0, nodes = 0, fails = 0, unk = 0, sweeps = 0, evals = 0, best = LPC_INF;
    int depth = 0, incomplete = 0;
    // the root is seeded by the caller's promise (or swept entirely); every other node differs from a fixpoint - its
    // parent's - in ONE variable, the one just branched on
    int seed_mode = n_root_seeds >= 0 ? 1 : 0, nseeds = n_root_seeds;
    const int* seeds = root_seeds;
    while(true) {
      bool bot;
      if constexpr(CD) {
        int node_sweeps = 0;
        bot = block_fixpoint_cd<HAS_DIV, TABLE_SMEM>(t, S, a_S, sx, sy, sz, sop, a_x, a_y, a_z, a_op, s_bot, s_evals, dmap, ng, ngp,
                                                     seed_mode, seeds, nseeds, 0, 1, node_sweeps);
        sweeps += node_sweeps;
        evals += *s_evals;
        seed_mode = 1; nseeds = 1; seeds = s_seed;
      }
      else {
        const long long before = sweeps;
        bot = block_fixpoint<HAS_DIV, TABLE_SMEM>(t, segs, S, a_S, sx, sy, sz, sop, a_x, a_y, a_z, a_op, s_bot, sweeps);
        evals += (sweeps - before) * ((npad + 31) / 32);
      }
      ++nodes;
      bool backtrack = false;
      if(bot) { ++fails; backtrack = true; }
      else {
        // first non-singleton branching variable
        if(tid == 0) *s_idx = 0x7fffffff;
        __syncthreads();
        for(int i = tid; i < nb; i += nthr) { const int2 d = S[bvars[i]]; if(d.x < d.y) { atomicMin(s_idx, i); break; } }
        __syncthreads();
        const int idx = *s_idx;
        if(idx == 0x7fffffff) {   // leaf: is_extractable
          int ok = 1;
          for(int i = tid; i < npad && ok; i += nthr) {
            const int2 a = S[sx[i]], bb = S[sy[i]], c = S[sz[i]];
            ok = ask_regs(sop[i], Itv(a.x, a.y), Itv(bb.x, bb.y), Itv(c.x, c.y));
          }
          if(__syncthreads_and(ok)) { ++sol; if(objective_var >= 0) best = min(best, (long long)S[objective_var].x); }
          else ++unk;
          backtrack = true;
        }
        else if(depth >= max_depth || (max_nodes > 0 && nodes >= max_nodes)) { incomplete = 1; break; }
        else {
          // snapshot of the right branch [mid + 1, ub] goes on the stack, the block continues with [lb, mid]
          const int v = bvars[idx];
          const int2 d = S[v];
          const int mid = (int)((long long)d.x + (((long long)d.y - (long long)d.x) >> 1));
          __syncthreads();                       // everyone has read d
          if(tid == 0) S[v] = make_int2(mid + 1, d.y);
          fence_async_smem();
          __syncthreads();
          if(tid == 0) {
            char* dst = (char*)(my_stack + (size_t)depth * store_stride);
            for(int o = 0; o < sbytes; o += 32768) bulk_s2g(dst + o, (char*)S + o, min(32768, sbytes - o));
            bulk_commit();
            bulk_wait_read0();                   // the copy engine has read the image; it may change again
            S[v] = make_int2(d.x, mid);
            my_vstack[depth] = v;
            *s_seed = v;
          }
          ++depth;
          a_maxd = max(a_maxd, depth);
          __syncthreads();
        }
      }
      if(backtrack) {
        if(depth == 0) break;
        --depth;
        fence_async_smem();                      // order this block's reads of S before the copy engine overwrites it
        __syncthreads();
        if(tid == 0) {                           // restore(): pop the most recent right branch
          bulk_wait0();                          // its write has completed
          mbar_expect_tx(&bars[0], (unsigned)sbytes);
          bulk_g2s_chunked((char*)S, (const char*)(my_stack + (size_t)depth * store_stride), sbytes, &bars[0]);
          *s_seed = my_vstack[depth];
        }
        mbar_wait(&bars[0], phase); phase ^= 1;
        __syncthreads();                         // s_seed is visible to the whole block
      }
    }
    if(tid == 0) {
      long long* ps = per_store + (size_t)cur * 6;
      ps[0] = sol; ps[1] = nodes; ps[2] = fails; ps[3] = best; ps[4] = incomplete; ps[5] = unk;
      a_sol += sol; a_nodes += nodes; a_fails += fails; a_unk += unk; a_inc += incomplete; a_sweeps += sweeps; a_evals += evals;
      a_best = min(a_best, best);
      int nx = atomicAdd(&ctl->next_store, 1);
      *s_next = nx < n_stores ? nx : -1;
    }
    __syncthreads();
    cur = *s_next;
    __syncthreads();
  }
  if(tid == 0) {
    bulk_wait0();
    atomicAdd((unsigned long long*)&ctl->n_solutions, (unsigned long long)a_sol);
    atomicAdd((unsigned long long*)&ctl->n_nodes, (unsigned long long)a_nodes);
    atomicAdd((unsigned long long*)&ctl->n_fails, (unsigned long long)a_fails);
    atomicAdd((unsigned long long*)&ctl->n_unknown_leaves, (unsigned long long)a_unk);
    atomicAdd((unsigned long long*)&ctl->n_incomplete, (unsigned long long)a_inc);
    atomicAdd((unsigned long long*)&ctl->sweeps_total, (unsigned long long)a_sweeps);
    atomicAdd((unsigned long long*)&ctl->deductions, (unsigned long long)(32 * a_evals));
    atomicMin(&ctl->best, a_best);
    atomicMax(&ctl->max_depth_seen, a_maxd);
  }
}

// EPS decomposition: store k := base with decision j halved according to bit j of (first_id + k).
__global__ void k_batch_init_split(int2* stores, int nvars, int n_stores, const int2* base, const int* dvars, int ndec,
                                   long long first_id, const long long* ids) {
  for(int k = blockIdx.x; k < n_stores; k += gridDim.x) {
    int2* S = stores + (size_t)k * nvars;
    for(int v = threadIdx.x; v < nvars; v += blockDim.x) S[v] = base[v];
    __syncthreads();
    const long long id = ids ? ids[k] : first_id + k;
    for(int j = threadIdx.x; j < ndec; j += blockDim.x) {
      const int v = dvars[j];
      const int2 d = base[v];
      const long long mid = (long long)d.x + (((long long)d.y - (long long)d.x) >> 1);
      S[v] = ((id >> j) & 1) ? make_int2((int)(mid + 1), d.y) : make_int2(d.x, (int)mid);
    }
    __syncthreads();
  }
}

// BatchCtl::payload from the reduction record (batch_internal.cuh), for the kernels that do not write it themselves.
__global__ void k_publish_payload(BatchCtl* ctl) {
  ctl->payload[0] = ctl->red[0]; ctl->payload[1] = ctl->red[1]; ctl->payload[2] = ctl->red[2];
  ctl->payload[3 + ctl->rank] = ctl->red[3];
}

} // namespace lpc

using namespace lpc;

typedef void (*batch_kernel_t)(TableDev, OpSegs, int2*, int, int, uint8_t*, int*, int*, BatchCtl*, int, int, int, const int*, int);

static batch_kernel_t pick_batch_kernel(bool has_div, bool table_smem, bool cd) {
  if(cd) {
    if(has_div) return table_smem ? k_pir_batch<true, true, true> : k_pir_batch<true, false, true>;
    return table_smem ? k_pir_batch<false, true, true> : k_pir_batch<false, false, true>;
  }
  if(has_div) return table_smem ? k_pir_batch<true, true, false> : k_pir_batch<true, false, false>;
  return table_smem ? k_pir_batch<false, true, false> : k_pir_batch<false, false, false>;
}
static size_t group_maps_bytes(const lpc_table* t) { return 3 * (size_t)(((t->dev.n_pad + 31) / 32 + 15) / 16 * 16); }

extern "C" {

int lpc_batch_create(const lpc_table* t, int32_t n_stores, lpc_batch** out) {
  LPC_REQUIRE(t && out, "null argument");
  LPC_REQUIRE(n_stores >= 0, "bad n_stores");
  LPC_REQUIRE(t->finalized, "lpc_table_finalize has not been called since the last change of the table");
  lpc_batch* b = new lpc_batch();
  b->table = t; b->n_stores = n_stores; b->nvars = t->dev.nvars; b->table_gen = t->generation;
  // ring slots are multiples of 16 B (bulk copy granularity); stores are packed at nvars*8 B in global memory, so
  // the per-store image must itself be a multiple of 16 B: require an even number of variables or pad by one.
  b->sbytes = (int)((((size_t)b->nvars * 8 + 15) / 16) * 16);
  if((size_t)b->sbytes != (size_t)b->nvars * 8) {
    delete b;
    set_error("lpc_batch_create: batched stores need an even number of variables (got %d); pad the model with one unused variable", t->dev.nvars);
    return LPC_ERR_UNSUPPORTED;
  }
  // every failure below destroys the half-built handle (and what it already allocated) before returning
  auto body = [&]() -> int {
    const size_t bytes = std::max<size_t>((size_t)n_stores * (size_t)b->nvars * 8, 16);
    LPC_CUDA(cudaMalloc((void**)&b->d, bytes));
    LPC_CUDA(cudaMalloc((void**)&b->d_flags, std::max(n_stores, 16)));
    LPC_CUDA(cudaMalloc((void**)&b->d_sweeps, std::max(n_stores, 4) * sizeof(int)));
    LPC_CUDA(cudaMalloc((void**)&b->d_obj, std::max(n_stores, 4) * sizeof(int)));
    LPC_CUDA(cudaMalloc((void**)&b->d_ctl, sizeof(BatchCtl)));
    LPC_CUDA(cudaMemset(b->d_ctl, 0, sizeof(BatchCtl)));
    LPC_CUDA(cudaMemset(b->d_flags, 0, std::max(n_stores, 16)));
    LPC_CUDA(cudaHostAlloc((void**)&b->h_ctl, sizeof(BatchCtl), cudaHostAllocDefault));
    memset(b->h_ctl, 0, sizeof(BatchCtl));
    LPC_CUDA(cudaHostAlloc((void**)&b->h_init, sizeof(BatchCtl), cudaHostAllocDefault));
    LPC_CUDA(cudaEventCreate(&b->ev0));
    LPC_CUDA(cudaEventCreate(&b->ev1));
    return LPC_OK;
  };
  const int rc_body = body();
  if(rc_body) { lpc_batch_destroy(b); return rc_body; }
  *out = b;
  return LPC_OK;
}

int lpc_batch_destroy(lpc_batch* b) {
  if(!b) return LPC_OK;
  cudaFree(b->d); cudaFree(b->d_flags); cudaFree(b->d_sweeps); cudaFree(b->d_obj); cudaFree(b->d_ctl); cudaFree(b->d_seeds);
  if(b->h_ctl) cudaFreeHost(b->h_ctl);
  if(b->h_init) cudaFreeHost(b->h_init);
  if(b->ev0) cudaEventDestroy(b->ev0);
  if(b->ev1) cudaEventDestroy(b->ev1);
  for(int c = 0; c < LPC_BATCH_CHUNKS; ++c) { if(b->e_in[c]) cudaEventDestroy(b->e_in[c]); if(b->e_k[c]) cudaEventDestroy(b->e_k[c]); }
  if(b->s_in) cudaStreamDestroy(b->s_in);
  if(b->s_k) cudaStreamDestroy(b->s_k);
  if(b->s_out) cudaStreamDestroy(b->s_out);
  cudaFree(b->d_ctl_chunk);
  if(b->h_ctl_chunk) cudaFreeHost(b->h_ctl_chunk);
  cudaFree(b->d_ptab); cudaFree(b->d_phdr); cudaFree(b->d_root); cudaFree(b->d_split_vars);
  delete b;
  return LPC_OK;
}

void* lpc_batch_device_ptr(lpc_batch* b) { return b ? (void*)b->d : nullptr; }
void* lpc_batch_reduction_device_ptr(lpc_batch* b) { return b ? (void*)b->d_ctl->red : nullptr; }

int lpc_batch_write(lpc_batch* b, int32_t first, int32_t n, const int32_t* lbub) {
  LPC_REQUIRE(b && (n == 0 || lbub), "null argument");
  LPC_REQUIRE(first >= 0 && n >= 0 && (long long)first + n <= b->n_stores, "range out of bounds");
  if(n) LPC_CUDA(cudaMemcpy(b->d + (size_t)first * b->nvars, lbub, (size_t)n * b->nvars * 8, cudaMemcpyHostToDevice));
  b->root_valid = false;   // arbitrary images: no common root is known any more
  b->split_fresh = false;
  return LPC_OK;
}

int lpc_batch_read(const lpc_batch* b, int32_t first, int32_t n, int32_t* lbub) {
  LPC_REQUIRE(b && (n == 0 || lbub), "null argument");
  LPC_REQUIRE(first >= 0 && n >= 0 && (long long)first + n <= b->n_stores, "range out of bounds");
  if(n) LPC_CUDA(cudaMemcpy(lbub, b->d + (size_t)first * b->nvars, (size_t)n * b->nvars * 8, cudaMemcpyDeviceToHost));
  return LPC_OK;
}

static int batch_init_split(lpc_batch* b, const int32_t* base_lbub, const int32_t* decision_vars, int32_t n_decisions,
                            int64_t first_id, const int64_t* ids);

int lpc_batch_init_split(lpc_batch* b, const int32_t* base_lbub, const int32_t* decision_vars, int32_t n_decisions,
                         int64_t first_id) {
  return batch_init_split(b, base_lbub, decision_vars, n_decisions, first_id, nullptr);
}

int lpc_batch_init_split_ids(lpc_batch* b, const int32_t* base_lbub, const int32_t* decision_vars, int32_t n_decisions,
                             const int64_t* ids) {
  LPC_REQUIRE(ids != nullptr || (b && b->n_stores == 0), "null ids");
  return batch_init_split(b, base_lbub, decision_vars, n_decisions, 0, ids);
}

static int batch_init_split(lpc_batch* b, const int32_t* base_lbub, const int32_t* decision_vars, int32_t n_decisions,
                            int64_t first_id, const int64_t* ids) {
  LPC_REQUIRE(b && base_lbub && (n_decisions == 0 || decision_vars), "null argument");
  LPC_REQUIRE(n_decisions >= 0 && n_decisions < 63, "bad n_decisions");
  for(int j = 0; j < n_decisions; ++j) LPC_REQUIRE(decision_vars[j] >= 0 && decision_vars[j] < b->nvars, "decision variable out of range");
  if(b->n_stores == 0) return LPC_OK;
  int2* d_base = nullptr;
  int* d_dec = nullptr;
  long long* d_ids = nullptr;
  auto body = [&]() -> int {
    LPC_CUDA(cudaMalloc((void**)&d_base, std::max<size_t>((size_t)b->nvars * 8, 16)));
    LPC_CUDA(cudaMalloc((void**)&d_dec, std::max(n_decisions, 1) * sizeof(int)));
    LPC_CUDA(cudaMemcpy(d_base, base_lbub, (size_t)b->nvars * 8, cudaMemcpyHostToDevice));
    if(n_decisions) LPC_CUDA(cudaMemcpy(d_dec, decision_vars, n_decisions * sizeof(int), cudaMemcpyHostToDevice));
    if(ids) {
      LPC_CUDA(cudaMalloc((void**)&d_ids, (size_t)b->n_stores * 8));
      LPC_CUDA(cudaMemcpy(d_ids, ids, (size_t)b->n_stores * 8, cudaMemcpyHostToDevice));
    }
    k_batch_init_split<<<std::min(b->n_stores, 148 * 16), 256>>>(b->d, b->nvars, b->n_stores, d_base, d_dec, n_decisions, first_id, d_ids);
    g_launches++;
    LPC_CUDA(cudaGetLastError());
    // every image of the batch is now a tightening of `base`: LPC_MODE_AUTO may drop the propagators entailed on it
    if(!b->d_root) LPC_CUDA(cudaMalloc((void**)&b->d_root, std::max<size_t>((size_t)b->nvars * 8, 16)));
    LPC_CUDA(cudaMemcpy(b->d_root, d_base, (size_t)b->nvars * 8, cudaMemcpyDeviceToDevice));
    b->root_valid = true;
    LPC_CUDA(cudaDeviceSynchronize());
    cudaFree(b->d_split_vars);   // the batch keeps the decision list
    b->d_split_vars = d_dec; d_dec = nullptr;
    b->n_split_vars = n_decisions; b->split_fresh = true;
    return LPC_OK;
  };
  const int rc = body();
  cudaFree(d_base);   // the scratch goes whether or not a step failed
  cudaFree(d_dec);
  cudaFree(d_ids);
  return rc;
}

int lpc_batch_set_seeds(lpc_batch* b, const int32_t* vars, int32_t n) {
  LPC_REQUIRE(b != nullptr && n >= -1 && (n <= 0 || vars), "bad argument");
  for(int i = 0; i < n; ++i) LPC_REQUIRE(vars[i] >= 0 && vars[i] < b->nvars, "seed variable out of range");
  cudaFree(b->d_seeds);
  b->d_seeds = nullptr;
  b->n_seeds = n;
  if(n > 0) {
    LPC_CUDA(cudaMalloc((void**)&b->d_seeds, (size_t)n * 4));
    LPC_CUDA(cudaMemcpy(b->d_seeds, vars, (size_t)n * 4, cudaMemcpyHostToDevice));
  }
  return LPC_OK;
}

// Launch the batch kernel on stores [first, first + count) of the batch with its own control block (plans are per handle).
static int batch_launch_range(lpc_batch* b, const lpc_fixpoint_opts* o, int32_t objective_var, int first, int count,
                              BatchCtl* d_ctl, BatchCtl* h_init, cudaStream_t st) {
  const lpc_table* t = b->table;
  LPC_REQUIRE(t->generation == b->table_gen, "the table changed since this batch was created: create a new batch");
  // LPC_MODE_SWEEP / LPC_MODE_AUTO: every sweep evaluates every record; LPC_MODE_WORKLIST: change-driven (block_fixpoint_cd),
  // seeded by lpc_batch_set_seeds when the caller made that promise. AUTO resolves to dense because that is what is faster
  // on the models measured: on config 4 one halved decision variable floods the 10k-record model within two sweeps (the
  // seeded change-driven run still evaluates 62 % of the dense run's propagators) and the flagging costs more than it saves
  // (33 vs 15.5 ms per 65,536 stores).
  const int cd = o->mode == LPC_MODE_WORKLIST ? 1 : 0;
  if(cd) { int rc = lpc_table_ensure_csr(const_cast<lpc_table*>(t)); if(rc) return rc; }
  if(!cd) {   // dense sweeps, large batch, model fits an SM: the grouped kernel over the packed table (pir_eps.cu)
    int used = 0;
    int rc = lpc_group_launch_resident(b, o, objective_var, first, count, d_ctl, h_init, st, &used);
    if(rc || used) return rc;
  }
  b->split_fresh = false;   // any fixpoint moves the images off the root of the split
  // launch plan: computed once per batch handle and mode (device attribute / occupancy queries are slow driver calls)
  if(!b->plan_ready[cd]) {
    int dev = 0, sms = 0, optin = 0;
    LPC_CUDA(cudaGetDevice(&dev));
    LPC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    LPC_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t max_smem = (size_t)optin;
    const size_t ring = 64 + 2 * (size_t)b->sbytes + (cd ? group_maps_bytes(t) : 0);
    const size_t tbl = (size_t)t->dev.n_pad * 13;
    if(ring > max_smem) {
      set_error("lpc_batch_fixpoint: a store of %d variables does not fit the shared-memory ring (%zu > %zu B); use lpc_fixpoint per store", b->nvars, ring, max_smem);
      return LPC_ERR_UNSUPPORTED;
    }
    // bulk copies need every sub-array 16-B aligned and sized: n_pad is a multiple of 16 (lpc_table_create)
    b->table_smem[cd] = ring + tbl <= max_smem;
    b->smem[cd] = b->table_smem[cd] ? ring + tbl : ring;
    batch_kernel_t kk = pick_batch_kernel(t->has_div, b->table_smem[cd], cd);
    LPC_CUDA(cudaFuncSetAttribute((const void*)kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem[cd]));
    b->threads[cd] = (int)std::min<long long>(1024, std::max<long long>(32, (t->dev.n_pad + 31) / 32 * 32));
    int per_sm = 0;
    LPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kk, b->threads[cd], b->smem[cd]));
    if(per_sm < 1 && b->threads[cd] > 256) {
      b->threads[cd] = 256;
      LPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kk, b->threads[cd], b->smem[cd]));
    }
    LPC_REQUIRE(per_sm > 0, "batch kernel does not fit on an SM");
    b->grid[cd] = std::max(1, std::min(b->n_stores, sms * per_sm));
    b->plan_ready[cd] = true;
  }
  batch_kernel_t k = pick_batch_kernel(t->has_div, b->table_smem[cd], cd);
  const int threads = b->threads[cd];
  const int grid = std::max(1, std::min(b->grid[cd], count));
  const size_t smem = b->smem[cd];
  memset(h_init, 0, sizeof(BatchCtl));
  h_init->red[3] = LPC_INF;
  h_init->next_store = grid;
  h_init->rank = b->rank; h_init->world = b->world;
  LPC_CUDA(cudaMemcpyAsync(d_ctl, h_init, sizeof(BatchCtl), cudaMemcpyHostToDevice, st));
  int2* dd = b->d + (size_t)first * b->nvars;
  uint8_t* fl = b->d_flags + first; int* sw = b->d_sweeps + first; int* ob = b->d_obj ? b->d_obj + first : nullptr;
  if(count > 0) {
    k<<<grid, threads, smem, st>>>(t->dev, t->opsegs, dd, count, b->sbytes, fl, sw, ob, d_ctl,
                                  objective_var, o->max_sweeps, o->stop_on_bot, b->d_seeds, b->n_seeds);
    g_launches++;
    LPC_CUDA(cudaGetLastError());
  }
  // these kernels leave the all-reduce payload to a one-thread epilogue (the grouped kernel's last block writes it itself)
  k_publish_payload<<<1, 1, 0, st>>>(d_ctl);
  g_launches++;
  LPC_CUDA(cudaGetLastError());
  return LPC_OK;
}

int lpc_batch_set_rank(lpc_batch* b, int32_t rank, int32_t world) {
  LPC_REQUIRE(b && world >= 1 && world <= LPC_MAX_RANKS && rank >= 0 && rank < world, "bad rank / world");
  b->rank = rank; b->world = world;
  return LPC_OK;
}

void* lpc_batch_payload_device_ptr(lpc_batch* b, int32_t* n_int64) {
  if(!b) return nullptr;
  if(n_int64) *n_int64 = 3 + b->world;
  return (void*)b->d_ctl->payload;
}

int lpc_batch_fixpoint_async(lpc_batch* b, const lpc_fixpoint_opts* o, int32_t objective_var) {
  LPC_REQUIRE(b != nullptr, "null batch");
  LPC_REQUIRE(objective_var < b->nvars, "objective variable out of range");
  LPC_REQUIRE(!b->pending, "a call is still in flight on this batch (collect it first)");
  { int rc = lpc_check_device(b->table->device, "lpc_batch_fixpoint"); if(rc) return rc; }
  lpc_fixpoint_opts def;
  if(!o) { lpc_fixpoint_default_opts(&def); o = &def; }
  cudaStream_t st = (cudaStream_t)o->stream;
  LPC_CUDA(cudaEventRecord(b->ev0, st));
  int rc = batch_launch_range(b, o, objective_var, 0, b->n_stores, b->d_ctl, b->h_init, st);
  if(rc) return rc;
  LPC_CUDA(cudaEventRecord(b->ev1, st));
  LPC_CUDA(cudaMemcpyAsync(b->h_ctl, b->d_ctl, sizeof(BatchCtl), cudaMemcpyDeviceToHost, st));
  b->last_stream = st;
  b->pending = true;
  b->n_chunks = 0;
  return LPC_OK;
}

int lpc_batch_collect(lpc_batch* b, lpc_batch_result* r) {
  LPC_REQUIRE(b != nullptr, "null batch");
  LPC_REQUIRE(b->pending, "no batch fixpoint in flight");
  LPC_CUDA(cudaStreamSynchronize(b->last_stream));
  if(b->n_chunks) {   // pipelined host call: the copy-out stream finishes last; fold the chunks' records into h_ctl
    LPC_CUDA(cudaStreamSynchronize(b->s_out));
    LPC_CUDA(cudaStreamSynchronize(b->s_k));
    BatchCtl acc;
    memset(&acc, 0, sizeof(acc));
    acc.red[3] = LPC_INF;
    for(int c = 0; c < b->n_chunks; ++c) {
      const BatchCtl& h = b->h_ctl_chunk[c];
      acc.red[0] += h.red[0]; acc.red[1] += h.red[1]; acc.red[2] += h.red[2];
      acc.red[3] = std::min(acc.red[3], h.red[3]);
      acc.sweeps_total += h.sweeps_total; acc.deductions += h.deductions;
      acc.max_sweeps_seen = std::max(acc.max_sweeps_seen, h.max_sweeps_seen);
      acc.hazard |= h.hazard;
    }
    *b->h_ctl = acc;
    // the device-side record stays consistent with a one-launch call (lpc_batch_reduction_device_ptr)
    LPC_CUDA(cudaMemcpy(b->d_ctl, b->h_ctl, sizeof(BatchCtl), cudaMemcpyHostToDevice));
    b->n_chunks = 0;
  }
  b->pending = false;
  if(r) {
    memset(r, 0, sizeof(*r));
    r->n_solution = b->h_ctl->red[0];
    r->n_bot = b->h_ctl->red[1];
    r->n_unknown = b->h_ctl->red[2];
    r->best_bound = (int32_t)b->h_ctl->red[3];
    r->max_sweeps_seen = b->h_ctl->max_sweeps_seen;
    r->sweeps_total = b->h_ctl->sweeps_total;
    r->deductions = b->h_ctl->deductions;
    r->overflow_hazard = b->h_ctl->hazard;
    float ms = 0;
    LPC_CUDA(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
    r->device_ms = ms;
  }
  return LPC_OK;
}

int lpc_batch_fixpoint(lpc_batch* b, const lpc_fixpoint_opts* o, int32_t objective_var, lpc_batch_result* r) {
  int rc = lpc_batch_fixpoint_async(b, o, objective_var);
  if(rc) return rc;
  return lpc_batch_collect(b, r);
}

int lpc_batch_fixpoint_host(lpc_batch* b, int32_t* lbub, const lpc_fixpoint_opts* o, int32_t objective_var,
                            lpc_batch_result* r) {
  LPC_REQUIRE(b && lbub, "null argument");
  LPC_REQUIRE(objective_var < b->nvars, "objective variable out of range");
  lpc_fixpoint_opts def;
  if(!o) { lpc_fixpoint_default_opts(&def); o = &def; }
  cudaStream_t st = (cudaStream_t)o->stream;
  b->root_valid = false;   // the images come from the caller: no common root is known
  b->split_fresh = false;
  const size_t store_bytes = (size_t)b->nvars * 8;
  const size_t bytes = (size_t)b->n_stores * store_bytes;
  int sms = 148;
  if(b->table) sms = std::max(1, b->table->sm_count);
  const char* ep = getenv("LPC_BATCH_PIPE");
  const bool pipe = (!ep || atoi(ep)) && b->n_stores >= 2 * LPC_BATCH_CHUNKS * 8 * sms && bytes >= (64u << 20);
  if(!pipe) {
    if(bytes) LPC_CUDA(cudaMemcpyAsync(b->d, lbub, bytes, cudaMemcpyHostToDevice, st));
    int rc = lpc_batch_fixpoint_async(b, o, objective_var);
    if(rc) return rc;
    if(bytes) LPC_CUDA(cudaMemcpyAsync(lbub, b->d, bytes, cudaMemcpyDeviceToHost, st));
    return lpc_batch_collect(b, r);
  }
  // Large batch from host memory: the link is full duplex and the stores are independent, so the batch goes through in
  // LPC_BATCH_CHUNKS chunks on three streams - chunk c + 1 is copied in and chunk c - 1 copied out while chunk c iterates.
  if(!b->s_in) {
    LPC_CUDA(cudaStreamCreateWithFlags(&b->s_in, cudaStreamNonBlocking));
    LPC_CUDA(cudaStreamCreateWithFlags(&b->s_k, cudaStreamNonBlocking));
    LPC_CUDA(cudaStreamCreateWithFlags(&b->s_out, cudaStreamNonBlocking));
    for(int c = 0; c < LPC_BATCH_CHUNKS; ++c) {
      LPC_CUDA(cudaEventCreateWithFlags(&b->e_in[c], cudaEventDisableTiming));
      LPC_CUDA(cudaEventCreateWithFlags(&b->e_k[c], cudaEventDisableTiming));
    }
    LPC_CUDA(cudaMalloc((void**)&b->d_ctl_chunk, LPC_BATCH_CHUNKS * sizeof(BatchCtl)));
    LPC_CUDA(cudaHostAlloc((void**)&b->h_ctl_chunk, 2 * LPC_BATCH_CHUNKS * sizeof(BatchCtl), cudaHostAllocDefault));
  }
  // order after whatever the caller queued on its stream
  LPC_CUDA(cudaEventRecord(b->ev0, st));
  LPC_CUDA(cudaStreamWaitEvent(b->s_in, b->ev0, 0));
  LPC_CUDA(cudaStreamWaitEvent(b->s_k, b->ev0, 0));
  LPC_CUDA(cudaStreamWaitEvent(b->s_out, b->ev0, 0));
  LPC_CUDA(cudaEventRecord(b->ev0, b->s_k));
  const int per = (b->n_stores + LPC_BATCH_CHUNKS - 1) / LPC_BATCH_CHUNKS;
  int nchunks = 0;
  for(int c = 0; c < LPC_BATCH_CHUNKS; ++c) {
    const int first = c * per, count = std::min(per, b->n_stores - first);
    if(count <= 0) break;
    char* hp = reinterpret_cast<char*>(lbub) + (size_t)first * store_bytes;
    char* dp = reinterpret_cast<char*>(b->d) + (size_t)first * store_bytes;
    LPC_CUDA(cudaMemcpyAsync(dp, hp, (size_t)count * store_bytes, cudaMemcpyHostToDevice, b->s_in));
    LPC_CUDA(cudaEventRecord(b->e_in[c], b->s_in));
    LPC_CUDA(cudaStreamWaitEvent(b->s_k, b->e_in[c], 0));
    int rc = batch_launch_range(b, o, objective_var, first, count, b->d_ctl_chunk + c, b->h_ctl_chunk + LPC_BATCH_CHUNKS + c, b->s_k);
    if(rc) return rc;
    LPC_CUDA(cudaMemcpyAsync(b->h_ctl_chunk + c, b->d_ctl_chunk + c, sizeof(BatchCtl), cudaMemcpyDeviceToHost, b->s_k));
    LPC_CUDA(cudaEventRecord(b->e_k[c], b->s_k));
    LPC_CUDA(cudaStreamWaitEvent(b->s_out, b->e_k[c], 0));
    LPC_CUDA(cudaMemcpyAsync(hp, dp, (size_t)count * store_bytes, cudaMemcpyDeviceToHost, b->s_out));
    ++nchunks;
  }
  LPC_CUDA(cudaEventRecord(b->ev1, b->s_k));
  b->n_chunks = nchunks;
  b->last_stream = b->s_k;
  b->pending = true;
  return lpc_batch_collect(b, r);
}

int lpc_batch_flags(const lpc_batch* b, uint8_t* out) {
  LPC_REQUIRE(b && (out || b->n_stores == 0), "null argument");
  if(b->n_stores) LPC_CUDA(cudaMemcpy(out, b->d_flags, b->n_stores, cudaMemcpyDeviceToHost));
  return LPC_OK;
}

int lpc_batch_search(lpc_batch* b, const int32_t* branch_vars, int32_t n_branch, const lpc_search_opts* o,
                     lpc_search_result* r, int64_t* per_store) {
  LPC_REQUIRE(b && (branch_vars || n_branch == 0) && n_branch >= 0, "bad argument");
  lpc_search_opts def;
  if(!o) { lpc_search_default_opts(&def); o = &def; }
  LPC_REQUIRE(o->objective_var < b->nvars, "objective variable out of range");
  LPC_REQUIRE(o->max_depth >= 1, "max_depth must be at least 1");
  for(int i = 0; i < n_branch; ++i) LPC_REQUIRE(branch_vars[i] >= 0 && branch_vars[i] < b->nvars, "branching variable out of range");
  const lpc_table* t = b->table;
  LPC_REQUIRE(t->generation == b->table_gen, "the table changed since this batch was created: create a new batch");
  cudaStream_t st = (cudaStream_t)o->stream;
  int dev = 0, sms = 0, optin = 0;
  LPC_CUDA(cudaGetDevice(&dev));
  LPC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  LPC_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  // change-driven nodes pay for their group maps on every node: a win on the 10,000-propagator model of config 4 (0.60 vs
  // 0.82 ms for 15.8 k nodes), a loss on a 500-propagator one (39 vs 18 ms for 1.1 M nodes), so the default goes by size
  const bool cd = o->change_driven < 0 ? t->dev.n >= 2048 : o->change_driven != 0;
  if(cd) { int rc = lpc_table_ensure_csr(const_cast<lpc_table*>(t)); if(rc) return rc; }
  const size_t base = 64 + (size_t)b->sbytes + (cd ? group_maps_bytes(t) : 0), tbl = (size_t)t->dev.n_pad * 13;
  if(base > (size_t)optin) { set_error("lpc_batch_search: a store of %d variables does not fit shared memory", b->nvars); return LPC_ERR_UNSUPPORTED; }
  const bool table_smem = base + tbl <= (size_t)optin;
  const size_t smem = table_smem ? base + tbl : base;
  typedef void (*search_kernel_t)(TableDev, OpSegs, const int2*, int, int, int2*, int*, int, const int*, int, int, long long, long long*,
                                  SearchCtl*, const int*, int);
  search_kernel_t k;
  if(cd) k = t->has_div ? (table_smem ? k_pir_search<true, true, true> : k_pir_search<true, false, true>)
                        : (table_smem ? k_pir_search<false, true, true> : k_pir_search<false, false, true>);
  else k = t->has_div ? (table_smem ? k_pir_search<true, true, false> : k_pir_search<true, false, false>)
                      : (table_smem ? k_pir_search<false, true, false> : k_pir_search<false, false, false>);
  LPC_CUDA(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // Threads per block. With plenty of subproblems a small model is searched faster by more, narrower blocks (four records
  // per thread: 200 variables / 500 propagators, 4,096 subproblems: 12.1 ms at 128 threads against 18.4 ms at 512 - a node's
  // cost is barriers and latency, not arithmetic); a batch that cannot fill the chip keeps one record per thread.
  const long long one_each = std::max<long long>(32, (t->dev.n_pad + 31) / 32 * 32);
  const long long four_each = std::max<long long>(128, (t->dev.n_pad / 4 + 31) / 32 * 32);
  const bool plenty = (long long)b->n_stores >= (long long)sms * std::max<long long>(1, 2048 / four_each);   // a full wave of them
  int threads = (int)std::min<long long>(t->has_div ? 256 : 1024, plenty ? four_each : one_each);
  if(const char* ev = getenv("LPC_SEARCH_TPB")) { const int v = atoi(ev); if(v >= 32 && v <= 1024 && v % 32 == 0) threads = v; }
  int per_sm = 0;
  LPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, smem));
  if(per_sm < 1 && threads > 256) { threads = 256; LPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, smem)); }
  LPC_REQUIRE(per_sm > 0, "search kernel does not fit on an SM");
  const int grid = std::max(1, std::min(b->n_stores, sms * per_sm));
  // scratch: snapshot stacks, branching order, per-store records, control block
  int2* d_stack = nullptr; int* d_bv = nullptr; long long* d_ps = nullptr; SearchCtl* d_ctl = nullptr; int* d_vstack = nullptr;
  SearchCtl h;
  memset(&h, 0, sizeof(h));
  h.best = LPC_INF; h.next_store = grid;
  auto body = [&]() -> int {
    LPC_CUDA(cudaMalloc((void**)&d_stack, std::max<size_t>((size_t)grid * o->max_depth * b->nvars * 8, 16)));
    LPC_CUDA(cudaMalloc((void**)&d_vstack, std::max<size_t>((size_t)grid * o->max_depth * 4, 16)));
    LPC_CUDA(cudaMalloc((void**)&d_bv, std::max<size_t>((size_t)n_branch * 4, 16)));
    LPC_CUDA(cudaMalloc((void**)&d_ps, std::max<size_t>((size_t)b->n_stores * 6 * 8, 16)));
    LPC_CUDA(cudaMalloc((void**)&d_ctl, sizeof(SearchCtl)));
    if(n_branch) LPC_CUDA(cudaMemcpyAsync(d_bv, branch_vars, (size_t)n_branch * 4, cudaMemcpyHostToDevice, st));
    LPC_CUDA(cudaMemcpyAsync(d_ctl, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    LPC_CUDA(cudaEventRecord(b->ev0, st));
    if(b->n_stores > 0) {
      k<<<grid, threads, smem, st>>>(t->dev, t->opsegs, b->d, b->n_stores, b->sbytes, d_stack, d_vstack, o->max_depth, d_bv, n_branch,
                                    o->objective_var, (long long)o->max_nodes, d_ps, d_ctl, b->d_seeds, b->n_seeds);
      g_launches++;
      LPC_CUDA(cudaGetLastError());
    }
    LPC_CUDA(cudaEventRecord(b->ev1, st));
    LPC_CUDA(cudaMemcpyAsync(&h, d_ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
    if(per_store && b->n_stores) LPC_CUDA(cudaMemcpyAsync(per_store, d_ps, (size_t)b->n_stores * 6 * 8, cudaMemcpyDeviceToHost, st));
    LPC_CUDA(cudaStreamSynchronize(st));
    return LPC_OK;
  };
  const int rc_body = body();
  cudaFree(d_stack); cudaFree(d_vstack); cudaFree(d_bv); cudaFree(d_ps); cudaFree(d_ctl);   // also when a step failed
  if(rc_body) return rc_body;
  if(r) {
    memset(r, 0, sizeof(*r));
    r->n_solutions = h.n_solutions; r->n_nodes = h.n_nodes; r->n_fails = h.n_fails; r->n_unknown_leaves = h.n_unknown_leaves;
    r->n_incomplete = h.n_incomplete; r->sweeps_total = h.sweeps_total; r->deductions = h.deductions;
    r->best_bound = (int32_t)h.best; r->max_depth_seen = h.max_depth_seen;
    float ms = 0;
    LPC_CUDA(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
    r->device_ms = ms;
  }
  return LPC_OK;
}

void lpc_search_default_opts(lpc_search_opts* o) {
  if(!o) return;
  memset(o, 0, sizeof(*o));
  o->max_nodes = 0; o->max_depth = 64; o->objective_var = -1;
  o->change_driven = -1;  // by table size (lpc_batch_search)
}

} // extern "C"
