// pir_eps.cu — batched mode, second generation: the grouped block kernel over a PACKED table, and the EPS-native call
// (sm_100a).
//
// What a solver does with the batched mode is embarrassingly parallel search (EPS): one root store, a list of decision
// variables, 2^k subproblem ids; it wants to know which subproblems fail, which are solutions, the best objective bound,
// and the stores of the few that survive - not 1 GiB of store images in each direction. So here
//   * the subproblem stores are GENERATED on the chip: a group copies the root image (L2-resident, 16 KB) into its
//     shared-memory slot with one bulk copy and halves the decision variables according to the bits of its id;
//   * only the stores that do NOT fail are written out, compacted (an atomic slot counter + one bulk store each), with
//     their subproblem index - the device-side form of `extract` (pir.hpp:890-898) after `is_extractable`
//     (pir.hpp:873-884);
//   * propagators that are already entailed on the ROOT store are dropped from the table before it is staged in shared
//     memory (`ask`, pir.hpp:417-438, used the way deinterpret(env, remove_entailed) uses it, pir.hpp:912-925): every
//     subproblem store is a tightening of the root, entailment survives tightening, and an entailed propagator can never
//     change a store again. The root of an EPS split is usually itself a fixpoint, where every
//     multiplication with fixed factors and every decided reified comparison is entailed - half of config 4's table.
// The table is staged as 8-byte packed records {x16 | y16 << 16, z16 | op << 16} (byte offsets into a store image), one
// run per operator, runs padded to an even length by repeating their last record (evaluating a propagator twice is
// harmless), so that a thread fetches TWO records with one 128-bit shared-memory load and has six gathers in flight
// before the first rule. The same kernel serves lpc_batch_fixpoint on resident store images (EPS = false).
#include "lpc_internal.cuh"
#include "smem_tma.cuh"
#include "batch_internal.cuh"

#include <algorithm>
#include <cstring>
#include <cstdlib>

namespace lpc {

#define LPC_PK_MAXRUN LPC_MAX_OPSEG
struct PackedHdr {
  int np;                          // packed records (even; padding duplicates included)
  int nruns;
  int start[LPC_PK_MAXRUN + 1];    // first packed record of each run (even); start[nruns] = np
  int op[LPC_PK_MAXRUN];           // device opcode of the run; -1 = mixed (opcode taken from each record)
  int n_live;                      // propagators kept (not entailed on the root)
  int n_total;                     // propagators of the table
  int root_flags;                  // EPS: bit0 = the root has an empty variable, bit1 = an infinite bound, bit2 = a finite
                                   // bound next to the int32 limits (lpc.h: overflow hazard); -1 = no root was scanned
};

// ---- table packer: one block; keeps table order inside every run ------------------------------------------------------
// hdr[0] / out: the packed table. hdr[1] / out1 (EPS, default mode): the records among those that mention a decision
// variable - the only propagators that can move anything in the FIRST sweep of a subproblem when the root is a common
// fixpoint of the table (checked here: no kept record moves a bound on the root). hdr[1].np = 0 when that does not hold,
// when there are more than cap1 such records, or when no decision list is given: the first sweep is then a full one.
constexpr int PK_K = 4;   // consecutive records per thread and iteration: their loads are in flight together
constexpr int PK_T = 512; // threads of the one block (4 x 512 records per iteration without spilling at 128 registers)
__global__ void __launch_bounds__(PK_T) k_pack_table(TableDev t, OpSegs segs, const int2* root, uint2* out, PackedHdr* hdr,
                                                     const int* dvars, int ndec, uint2* out1, int cap1, const int2* scan_root) {
  __shared__ unsigned s_warp[PK_T / 32];    // per warp: kept records | first-sweep records << 16
  __shared__ int s_rflags;
  __shared__ unsigned s_dmap[256];   // bit v: variable v is a decision variable (the grouped kernel takes nvars <= 8191)
  __shared__ int s_moves;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool want1 = dvars != nullptr && root != nullptr && out1 != nullptr && cap1 >= 2 && t.nvars <= 8192;
  if(tid < 256) s_dmap[tid] = 0;
  if(tid == 0) { s_moves = 0; s_rflags = 0; }
  // both headers travel to shared memory as whole structs: the entries behind the last run are zero, not whatever the
  // allocation held
  for(int i = tid; i < (int)(2 * sizeof(PackedHdr) / 4); i += PK_T) reinterpret_cast<int*>(hdr)[i] = 0;
  __syncthreads();
  if(scan_root) {   // what every subproblem store inherits from the root (halving a variable keeps its bounds inside the old ones)
    int rf = 0;
    for(int v = tid; v < t.nvars; v += PK_T) {
      const int2 d = scan_root[v];
      rf |= (d.x > d.y ? 1 : 0) | ((d.x == LPC_MINF || d.y == LPC_INF) ? 2 : 0) | ((near_inf_lo(d.x) | near_inf_hi(d.y)) ? 4 : 0);
    }
    if(rf) atomicOr(&s_rflags, rf);
  }
  if(want1) for(int j = tid; j < ndec; j += PK_T) atomicOr(&s_dmap[dvars[j] >> 5], 1u << (dvars[j] & 31));
  __syncthreads();
  const int nseg = segs.n == 0 ? 1 : segs.n;
  int base = 0, nruns = 0, n_live = 0;
  int base1 = 0, nruns1 = 0;
  bool over1 = false;
  int moves = 0;
  for(int r = 0; r < nseg; ++r) {
    const int s0 = segs.n == 0 ? 0 : segs.start[r], s1 = segs.n == 0 ? (int)t.n : segs.start[r + 1];
    const int run_op = segs.n == 0 ? -1 : (int)segs.op[r];
    const int run_begin = base, run_begin1 = base1;
    for(int i0 = s0; i0 < s1; i0 += PK_T * PK_K) {
      int op[PK_K], x[PK_K], y[PK_K], z[PK_K];
      int2 a[PK_K], b[PK_K], c[PK_K];
#pragma unroll
      for(int k = 0; k < PK_K; ++k) {
        const int i = i0 + PK_K * tid + k;
        const bool in = i < s1;
        op[k] = in ? (int)t.op[i] : D_NOP; x[k] = in ? t.x[i] : 0; y[k] = in ? t.y[i] : 0; z[k] = in ? t.z[i] : 0;
      }
      if(root) {
#pragma unroll
        for(int k = 0; k < PK_K; ++k) { a[k] = root[x[k]]; b[k] = root[y[k]]; c[k] = root[z[k]]; }
      }
      unsigned livem = 0, incm = 0;
#pragma unroll
      for(int k = 0; k < PK_K; ++k) {
        if(op[k] == D_NOP) continue;
        bool live = true, inc = false;
        if(root) {
          // an empty operand: keep the record (the store is at bot anyway, nothing is entailed on bot)
          const bool any_bot = (a[k].x > a[k].y) | (b[k].x > b[k].y) | (c[k].x > c[k].y);
          if(!any_bot && ask_regs(op[k], Itv(a[k].x, a[k].y), Itv(b[k].x, b[k].y), Itv(c[k].x, c[k].y))) live = false;
          else if(want1) {
            Itv r1(a[k].x, a[k].y), r2(b[k].x, b[k].y), r3(c[k].x, c[k].y);
            deduce_regs<true>(op[k], r1, r2, r3);
            moves |= any_bot | (r1.lb != a[k].x) | (r1.ub != a[k].y) | (r2.lb != b[k].x) | (r2.ub != b[k].y) | (r3.lb != c[k].x) | (r3.ub != c[k].y);
            inc = ((s_dmap[x[k] >> 5] >> (x[k] & 31)) | (s_dmap[y[k] >> 5] >> (y[k] & 31)) | (s_dmap[z[k] >> 5] >> (z[k] & 31))) & 1;
          }
        }
        livem |= (live ? 1u : 0u) << k;
        incm |= (inc ? 1u : 0u) << k;
      }
      // exclusive positions: inclusive scan over the warp of (kept | first-sweep << 16), then over the warps
      const unsigned mine = (unsigned)__popc(livem) | ((unsigned)__popc(incm) << 16);
      unsigned incl = mine;
#pragma unroll
      for(int off = 1; off < 32; off <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, incl, off); if(lane >= off) incl += u; }
      if(lane == 31) s_warp[warp] = incl;
      __syncthreads();
      unsigned before = 0, total = 0;
      for(int w = 0; w < PK_T / 32; ++w) { const unsigned cw = s_warp[w]; if(w < warp) before += cw; total += cw; }
      const int total0 = (int)(total & 0xffffu), total1 = (int)(total >> 16);
      int pos = base + (int)((before + incl - mine) & 0xffffu), pos1 = base1 + (int)((before + incl - mine) >> 16);
      const bool fits1 = !over1 && base1 + total1 + 1 <= cap1;
#pragma unroll
      for(int k = 0; k < PK_K; ++k) {
        const uint2 rec = make_uint2((unsigned)(8 * x[k]) | ((unsigned)(8 * y[k]) << 16), (unsigned)(8 * z[k]) | ((unsigned)op[k] << 16));
        if((livem >> k) & 1) out[pos++] = rec;
        if(fits1 && ((incm >> k) & 1)) out1[pos1++] = rec;
      }
      if(!fits1 && total1 > 0) over1 = true;
      base += total0;
      if(!over1) base1 += total1;
      __syncthreads();
    }
    n_live += base - run_begin;
    if((base - run_begin) & 1) {   // odd run: repeat its last record
      if(tid == 0) out[base] = out[base - 1];
      ++base;
    }
    if(!over1 && ((base1 - run_begin1) & 1)) {
      if(tid == 0) out1[base1] = out1[base1 - 1];
      ++base1;
    }
    __syncthreads();
    if(base > run_begin) {
      if(tid == 0) { hdr[0].start[nruns] = run_begin; hdr[0].op[nruns] = run_op; }
      ++nruns;
    }
    if(want1 && !over1 && base1 > run_begin1) {
      if(tid == 0) { hdr[1].start[nruns1] = run_begin1; hdr[1].op[nruns1] = run_op; }
      ++nruns1;
    }
  }
  if(moves) s_moves = 1;
  __syncthreads();
  if(tid == 0) {
    hdr[0].start[nruns] = base;
    hdr[0].np = base; hdr[0].nruns = nruns; hdr[0].n_live = n_live; hdr[0].n_total = (int)t.n;
    hdr[0].root_flags = scan_root ? s_rflags : -1;
    const bool ok1 = want1 && !over1 && s_moves == 0;
    hdr[1].start[ok1 ? nruns1 : 0] = ok1 ? base1 : 0;
    hdr[1].np = ok1 ? base1 : 0; hdr[1].nruns = ok1 ? nruns1 : 0; hdr[1].n_live = ok1 ? base1 : 0; hdr[1].n_total = (int)t.n;
    hdr[1].root_flags = -1;
  }
}

// ---- the exchange of the reduction record between the GPUs of one node, fused into the batch kernel ------------------------
// Every rank owns an inbox of 2 x LPC_MAX_RANKS slots in its device memory, opened by every peer through CUDA IPC
// (lpc_eps_peer_export / _connect). The block that finishes a rank's batch last writes the rank's record into slot
// [epoch & 1][my rank] of EVERY inbox (its own included) - 40 bytes per peer over NVLink, straight from the kernel -, then
// releases the slot's epoch word at system scope. k_eps_fold, a one-thread kernel queued behind the batch kernel, acquires
// the epoch words of its own inbox and folds the records into the payload. No collective library call on the path. Two
// slot sets alternate: a rank can run at most one step ahead of a peer (its fold of step k needs every peer's record of
// step k), so the records of step k + 1 never land on records of step k that are still unread.
struct PeerSlot { long long rec[4]; long long epoch; long long pad[3]; };
static_assert(sizeof(PeerSlot) == 64, "one slot per 64-byte line");

__device__ __forceinline__ void st_release_sys(long long* p, long long v) {
  asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long ld_acquire_sys(const long long* p) {
  long long v;
  asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// payload[0..2] = sum of the ranks' counters, payload[3 + r] = rank r's best bound: the layout the all-reduce produces.
// A peer that never arrives (2^28 polls, a few seconds) leaves payload[0] = -1 instead of hanging the GPU.
__global__ void k_eps_fold(const PeerSlot* inbox, int world, long long epoch, BatchCtl* ctl) {
  long long sum[3] = {0, 0, 0};
  bool ok = true;
  for(int r = 0; r < world && ok; ++r) {
    const PeerSlot* s = inbox + (epoch & 1) * LPC_MAX_RANKS + r;
    long long polls = 0;
    while(ld_acquire_sys(&s->epoch) != epoch) { if(++polls > (1ll << 28)) { ok = false; break; } }
    if(!ok) break;
    for(int i = 0; i < 3; ++i) sum[i] += s->rec[i];
    ctl->payload[3 + r] = s->rec[3];
  }
  for(int i = 0; i < 3; ++i) ctl->payload[i] = ok ? sum[i] : -1;
}

// ---- the grouped kernel ---------------------------------------------------------------------------------------------
struct GroupArgs {
  const uint2* ptab; const PackedHdr* hdr;   // hdr[0]: the packed table; hdr[1] / ptab1: the first-sweep table (k_pack_table)
  const uint2* ptab1;              // null: every sweep is a full one
  int2* stores;                    // !EPS: resident images [n_stores][nvars], fixpoints written back in place
  const int2* root;                // EPS: the root store
  const int* dvars; int ndec;      // EPS: decision variables (bit j of the id halves dvars[j])
  const long long* ids;            // EPS: subproblem ids (null: first_id + k)
  long long first_id;
  int n_stores, nvars, sbytes;
  uint8_t* flags; int* sweeps_out; int* obj_out;
  int2* surv; int* surv_idx; int surv_cap;   // EPS: compacted non-failed stores and their subproblem index k
  BatchCtl* ctl;
  int objective_var, max_sweeps, stop_on_bot;
  PeerSlot* const* peers;          // multi-GPU: the inboxes of all ranks (device array of `world` pointers), or null
  int my_rank, world;
  long long epoch;
};

__device__ __forceinline__ uint4 lds_v4(unsigned addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
// Named barriers of a thread group. `bar.sync` / `bar.red` are the .aligned forms: every thread of a warp must execute them
// together, so a warp that may still be split by an earlier data-dependent branch (the per-lane exit of the ask loop, a
// predicated atomic) is brought back together first (compute-sanitizer synccheck flags it otherwise).
__device__ __forceinline__ void gbar_sync(int id, int n) {
  __syncwarp();
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ int gbar_or(int id, int n, int pred) {
  int r;
  __syncwarp();
  asm volatile("{ .reg .pred p, q; setp.ne.s32 q, %3, 0; bar.red.or.pred p, %1, %2, q; selp.s32 %0, 1, 0, p; }"
               : "=r"(r) : "r"(id), "r"(n), "r"(pred) : "memory");
  return r;
}
__device__ __forceinline__ int gbar_and(int id, int n, int pred) {
  int r;
  __syncwarp();
  asm volatile("{ .reg .pred p, q; setp.ne.s32 q, %3, 0; bar.red.and.pred p, %1, %2, q; selp.s32 %0, 1, 0, p; }"
               : "=r"(r) : "r"(id), "r"(n), "r"(pred) : "memory");
  return r;
}

// An emptied variable is signalled to the whole group at once (a predicated store to the group's bot word), so that the
// other threads can abandon the rest of the sweep: the store has failed, whatever else the sweep computes is moot.
__device__ __forceinline__ void signal_bot(unsigned a_bot, int bn) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %1, 0; @q st.shared.s32 [%0], %1; }" ::"r"(a_bot), "r"(bn) : "memory");
}

// One propagator on gathered bounds. FIN: every bound of the store was finite when it was loaded; bounds only tighten,
// so the infinity guards of pir.hpp:759-764 can never fire and `x = y + z` is six fused add-min/max.
// The join: a lane whose record moved anything issues all six shared-memory reductions (the ones of bounds that did not
// move are no-ops of the lattice join), and looks for an emptied operand there and only there - an operand that was
// empty before the first sweep is found by the scan at load time, and a bound can only cross its partner by moving.
// JOIN = 1 (default) / 2: every reduction individually guarded by "this bound moved", inside the moved-region (1) or
// without a region at all (2). ptxas turns the guard into a branch around the ATOMS (a shared-memory atomic cannot be
// predicated), so this costs instructions - but the kernel is bound by the shared-memory pipe (ncu: l1tex throughput 94 %,
// half of the wavefronts bank conflicts, issue slots 49 % busy), and a no-op reduction occupies its banks like a real one.
__device__ __forceinline__ void reds_max_if_gt(unsigned addr, int nw, int old) {
  asm volatile("{ .reg .pred q; setp.gt.s32 q, %1, %2; @q red.shared.max.s32 [%0], %1; }" ::"r"(addr), "r"(nw), "r"(old) : "memory");
}
__device__ __forceinline__ void reds_min_if_lt(unsigned addr, int nw, int old) {
  asm volatile("{ .reg .pred q; setp.lt.s32 q, %1, %2; @q red.shared.min.s32 [%0], %1; }" ::"r"(addr), "r"(nw), "r"(old) : "memory");
}

// JOIN = 3: no atomics at all - a bound that moved is written with a plain predicated store. Two lanes that tighten the
// same bound in the same sweep may overwrite each other (the weaker value can win, a bound can even step back within a
// sweep), which the fixpoint tolerates: every value ever written is the result of a sound rule on values that contain the
// greatest fixpoint G, so it contains G; a written value is strictly tighter than a value read in the same sweep, hence
// (induction over the writes of a sweep) than the bound at the sweep's start, so every sweep that writes makes strict
// progress and the loop ends; and the sweep that ends it wrote nothing, i.e. every propagator was evaluated on the
// final store and moved nothing: a common fixpoint inside the initial store, which is below G. So the result is G, bit
// for bit, as with atomic joins - without the shared-memory atomics' bank and same-address serialisation.
__device__ __forceinline__ void sts_if_gt(unsigned addr, int nw, int old) {
  asm volatile("{ .reg .pred q; setp.gt.s32 q, %1, %2; @q st.shared.s32 [%0], %1; }" ::"r"(addr), "r"(nw), "r"(old) : "memory");
}
__device__ __forceinline__ void sts_if_lt(unsigned addr, int nw, int old) {
  asm volatile("{ .reg .pred q; setp.lt.s32 q, %1, %2; @q st.shared.s32 [%0], %1; }" ::"r"(addr), "r"(nw), "r"(old) : "memory");
}

template <int OP, bool HAS_DIV, bool FIN, int JOIN>
__device__ __forceinline__ void pk_rule(int op, int2 a, int2 b, int2 c, unsigned ax, unsigned ay, unsigned az,
                                        unsigned& macc, int& bacc, unsigned a_bot) {
  Itv r1(a.x, a.y), r2(b.x, b.y), r3(c.x, c.y);
  if(FIN && OP == D_ADD) {
    // pir.hpp:759-764; later lines see the bounds the earlier ones just tightened, as in the reference
    r1.lb = max(r1.lb, wadd(r2.lb, r3.lb)); r1.ub = min(r1.ub, wadd(r2.ub, r3.ub));
    r2.lb = max(r2.lb, wsub(r1.lb, r3.ub)); r2.ub = min(r2.ub, wsub(r1.ub, r3.lb));
    r3.lb = max(r3.lb, wsub(r1.lb, r2.ub)); r3.ub = min(r3.ub, wsub(r1.ub, r2.lb));
  }
  else deduce_regs<HAS_DIV>(op, r1, r2, r3);
  const unsigned moved = (unsigned)((r1.lb ^ a.x) | (r1.ub ^ a.y) | (r2.lb ^ b.x) | (r2.ub ^ b.y) | (r3.lb ^ c.x) | (r3.ub ^ c.y));
  if(JOIN == 3) {
    sts_if_gt(ax, r1.lb, a.x); sts_if_lt(ax + 4, r1.ub, a.y);
    sts_if_gt(ay, r2.lb, b.x); sts_if_lt(ay + 4, r2.ub, b.y);
    sts_if_gt(az, r3.lb, c.x); sts_if_lt(az + 4, r3.ub, c.y);
    { const int bn = (r1.lb > r1.ub) | (r2.lb > r2.ub) | (r3.lb > r3.ub); signal_bot(a_bot, bn); bacc |= bn; }
  }
  else if(JOIN == 2) {
    reds_max_if_gt(ax, r1.lb, a.x); reds_min_if_lt(ax + 4, r1.ub, a.y);
    reds_max_if_gt(ay, r2.lb, b.x); reds_min_if_lt(ay + 4, r2.ub, b.y);
    reds_max_if_gt(az, r3.lb, c.x); reds_min_if_lt(az + 4, r3.ub, c.y);
    { const int bn = (r1.lb > r1.ub) | (r2.lb > r2.ub) | (r3.lb > r3.ub); signal_bot(a_bot, bn); bacc |= bn; }
  }
  else if(moved) {
    if(JOIN == 4) {   // the plain stores of JOIN = 3, inside the moved-region
      sts_if_gt(ax, r1.lb, a.x); sts_if_lt(ax + 4, r1.ub, a.y);
      sts_if_gt(ay, r2.lb, b.x); sts_if_lt(ay + 4, r2.ub, b.y);
      sts_if_gt(az, r3.lb, c.x); sts_if_lt(az + 4, r3.ub, c.y);
    }
    else if(JOIN == 1) {
      reds_max_if_gt(ax, r1.lb, a.x); reds_min_if_lt(ax + 4, r1.ub, a.y);
      reds_max_if_gt(ay, r2.lb, b.x); reds_min_if_lt(ay + 4, r2.ub, b.y);
      reds_max_if_gt(az, r3.lb, c.x); reds_min_if_lt(az + 4, r3.ub, c.y);
    }
    else {
      reds_max(ax, r1.lb); reds_min(ax + 4, r1.ub);
      reds_max(ay, r2.lb); reds_min(ay + 4, r2.ub);
      reds_max(az, r3.lb); reds_min(az + 4, r3.ub);
    }
    { const int bn = (r1.lb > r1.ub) | (r2.lb > r2.ub) | (r3.lb > r3.ub); signal_bot(a_bot, bn); bacc |= bn; }
  }
  macc |= moved;
}

// The pairs [p0, p1) of one run: two records per thread and iteration, in chunks of four iterations after each of which
// the group's bot word is looked at (stop_on_bot): most EPS subproblems fail, and a failed store's sweep is cut short as
// soon as any thread has emptied a variable. `nev` counts the records this thread evaluated.
template <int OP, bool HAS_DIV, bool FIN, int JOIN>
__device__ __forceinline__ int pk_sweep_run(int p0, int p1, unsigned a_T, unsigned a_S, int tid, int nthr, unsigned a_bot, int stop,
                                            unsigned& nev) {
  unsigned macc = 0;
  int bacc = 0;
  for(int c0 = p0; c0 < p1; c0 += 4 * nthr) {
    const int c1 = min(c0 + 4 * nthr, p1);
    for(int p = c0 + tid; p < c1; p += nthr) {
      const uint4 q = lds_v4(a_T + 16u * (unsigned)p);
      const unsigned ax0 = a_S + (q.x & 0xffffu), ay0 = a_S + (q.x >> 16), az0 = a_S + (q.y & 0xffffu);
      const unsigned ax1 = a_S + (q.z & 0xffffu), ay1 = a_S + (q.z >> 16), az1 = a_S + (q.w & 0xffffu);
      const int2 a0 = lds_itv(ax0), b0 = lds_itv(ay0), c0v = lds_itv(az0);
      const int2 a1 = lds_itv(ax1), b1 = lds_itv(ay1), c1v = lds_itv(az1);
      pk_rule<OP, HAS_DIV, FIN, JOIN>(OP < 0 ? (int)(q.y >> 16) : OP, a0, b0, c0v, ax0, ay0, az0, macc, bacc, a_bot);
      pk_rule<OP, HAS_DIV, FIN, JOIN>(OP < 0 ? (int)(q.w >> 16) : OP, a1, b1, c1v, ax1, ay1, az1, macc, bacc, a_bot);
    }
    nev += 2u * (unsigned)max(0, (c1 - c0 - tid + nthr - 1) / nthr);
    if(stop && lds_s32(a_bot)) break;
  }
  return (macc != 0u ? 1 : 0) | (bacc ? 2 : 0);
}

template <bool HAS_DIV, int JOIN>
__device__ __forceinline__ int pk_sweep(const PackedHdr& h, unsigned a_T, unsigned a_S, int tid, int nthr, bool fin, unsigned a_bot,
                                        int stop, unsigned& nev) {
  int f = 0;
  for(int r = 0; r < h.nruns; ++r) {
    if(r && stop && lds_s32(a_bot)) break;
    const int p0 = h.start[r] >> 1, p1 = h.start[r + 1] >> 1;
#define LPC_RUN(O) case O: f |= pk_sweep_run<O, HAS_DIV, false, JOIN>(p0, p1, a_T, a_S, tid, nthr, a_bot, stop, nev); break;
    switch(h.op[r]) {
      case D_ADD: f |= fin ? pk_sweep_run<D_ADD, HAS_DIV, true, JOIN>(p0, p1, a_T, a_S, tid, nthr, a_bot, stop, nev)
                           : pk_sweep_run<D_ADD, HAS_DIV, false, JOIN>(p0, p1, a_T, a_S, tid, nthr, a_bot, stop, nev); break;
      LPC_RUN(D_MUL) LPC_RUN(D_MIN) LPC_RUN(D_MAX) LPC_RUN(D_EQ) LPC_RUN(D_LEQ)
      default: f |= pk_sweep_run<-1, HAS_DIV, false, JOIN>(p0, p1, a_T, a_S, tid, nthr, a_bot, stop, nev); break;
    }
#undef LPC_RUN
  }
  return f;
}

// G groups of 1024 / G threads per block, one block per SM; each group owns one shared-memory store slot, claims its
// own stores from a global counter and synchronises on its own named barrier, so one group's barriers and copy waits are
// filled with the other groups' instructions. Shared memory: [0, 128) mbarriers | [128, 640) per-group scalars |
// [640, 1024) header | G store slots | packed table.
template <bool HAS_DIV, int G, bool EPS, int JOIN = 1>
__global__ void __launch_bounds__(1024, 1) k_pir_group(GroupArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int nthr = 1024 / G;
  const int grp = threadIdx.x / nthr, tid = threadIdx.x % nthr, bid = 1 + grp;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem);   // [grp]: store slot, [G]: table
  // per group 64 B of scalars: the scheduler's next store, the bot word, the records evaluated on the current store, and
  // the group's running totals (kept here rather than in thread 0's registers: they would be live in every thread)
  struct GroupAcc { int next; int bot; unsigned nev; int best; long long sol, nbot, unk, sweeps, ded; int maxsw; int bot1; };
  static_assert(sizeof(GroupAcc) <= 64, "group scalars");
  GroupAcc* ga = reinterpret_cast<GroupAcc*>(smem + 128 + 64 * grp);
  int* s_next = &ga->next;
  // Two bot words, one per sweep parity: a thread that has left the sweep-ending barrier and runs ahead into the next
  // sweep must not change what a slower thread is about to read as the outcome of the sweep just ended (else the two
  // disagree on whether the store has failed and part ways on the group's barrier)
  volatile int* s_bot = &ga->bot;
  volatile int* s_bot1 = &ga->bot1;
  unsigned* s_nev = &ga->nev;
  PackedHdr* sh = reinterpret_cast<PackedHdr*>(smem + 640);
  PackedHdr* sh1 = reinterpret_cast<PackedHdr*>(smem + 640 + 192);
  static_assert(sizeof(PackedHdr) <= 192, "two headers share the 384-byte header area");
  const int sbytes = A.sbytes;
  int2* S = reinterpret_cast<int2*>(smem + 1024 + (size_t)grp * sbytes);
  char* tb = reinterpret_cast<char*>(smem + 1024 + (size_t)G * sbytes);
  const size_t store_stride = (size_t)A.nvars;
  unsigned long long* tbar = &bars[G];
  const int np = A.hdr->np;
  // > 0: sweep 1 of every store runs on the first-sweep table (not under a sweep budget: a caller who asks for k sweeps gets
  // k full ones)
  const int np1 = (A.ptab1 && A.max_sweeps == 0) ? A.hdr[1].np : 0;

  int* s_groups_done = reinterpret_cast<int*>(smem + 96);   // behind the G + 1 mbarriers of the first 128 bytes
  static_assert((G + 1) * 8 <= 96, "mbarriers overlap the group counter");
  if(threadIdx.x == 0) {
    for(int i = 0; i <= G; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
    *s_groups_done = 0;
  }
  if(threadIdx.x < (int)(sizeof(PackedHdr) / 4)) {
    reinterpret_cast<int*>(sh)[threadIdx.x] = reinterpret_cast<const int*>(A.hdr)[threadIdx.x];
    if(np1 > 0) reinterpret_cast<int*>(sh1)[threadIdx.x] = reinterpret_cast<const int*>(A.hdr + 1)[threadIdx.x];
  }
  __syncthreads();
  int cur = G * blockIdx.x + grp < A.n_stores ? G * blockIdx.x + grp : -1;
  if(threadIdx.x == 0 && np > 0) {
    mbar_expect_tx(tbar, (unsigned)((np + np1) * 8));
    bulk_g2s_chunked(tb, (const char*)A.ptab, (unsigned)(np * 8), tbar);
    if(np1 > 0) bulk_g2s_chunked(tb + (size_t)np * 8, (const char*)A.ptab1, (unsigned)(np1 * 8), tbar);
  }
  if(tid == 0 && cur >= 0) {
    mbar_expect_tx(&bars[grp], (unsigned)sbytes);
    bulk_g2s_chunked((char*)S, EPS ? (const char*)A.root : (const char*)(A.stores + cur * store_stride), sbytes, &bars[grp]);
  }
  if(np > 0) mbar_wait(tbar, 0);
  const unsigned a_T = smem_u32(tb), a_S = smem_u32(S), a_T1 = a_T + 8u * (unsigned)np;

  if(tid == 0) { ga->best = LPC_INF; ga->sol = ga->nbot = ga->unk = ga->sweeps = ga->ded = 0; ga->maxsw = 0; }
  unsigned phase = 0;
  const unsigned a_sbot0 = smem_u32(const_cast<int*>(s_bot)), a_sbot1 = smem_u32(const_cast<int*>(s_bot1));
  while(cur >= 0) {
    if(tid == 0) {   // claim the next store of this group
      int nx = atomicAdd(&A.ctl->next_store, 1);
      if(nx >= A.n_stores) nx = -1;
      *s_next = nx;
      *s_bot = 0;
      *s_bot1 = 0;
      *s_nev = 0;
    }
    mbar_wait(&bars[grp], phase);
    phase ^= 1;
    int f0 = 0, inf = 0, hz = 0;
    bool bot, fin;
    const int rflags = EPS ? sh->root_flags : -1;
    if(EPS) {   // the subproblem: bit j of the id keeps the lower (0) or the upper (1) half of decision variable j
      const long long id = A.ids ? A.ids[cur] : A.first_id + cur;
      for(int j = tid; j < A.ndec; j += nthr) {
        const int v = A.dvars[j];
        const int2 d = S[v];
        const long long mid = (long long)d.x + (((long long)d.y - (long long)d.x) >> 1);
        const int2 h = ((id >> j) & 1) ? make_int2((int)(mid + 1), d.y) : make_int2(d.x, (int)mid);
        S[v] = h;
        f0 |= h.x > h.y;   // the upper half of a singleton is empty
      }
    }
    if(EPS && rflags >= 0) {
      // Everything but the halved variables is the root, whose emptiness / finiteness / overflow hazard k_pack_table has
      // scanned once for the whole batch; a half of a variable lies inside its old bounds, so it adds neither an infinite
      // bound nor a hazard - only, for a singleton, an empty upper half.
      bot = (gbar_or(bid, nthr, f0) != 0) | ((rflags & 1) != 0);   // also orders thread 0's s_bot / s_next writes before their readers
      fin = (rflags & 2) == 0;
      hz = (rflags & 4) != 0 && tid == 0;   // one thread reports it
    }
    else {
      if(EPS) gbar_sync(bid, nthr);
      for(int v = tid; v < A.nvars; v += nthr) {
        const int2 d = S[v];
        f0 |= d.x > d.y;
        inf |= (d.x == LPC_MINF) | (d.y == LPC_INF);
        hz |= near_inf_lo(d.x) | near_inf_hi(d.y);
      }
      bot = gbar_or(bid, nthr, f0) != 0;   // also orders thread 0's s_bot / s_next writes before their readers
      fin = gbar_or(bid, nthr, inf) == 0;
    }
    int sweeps = 0;
    unsigned nev = 0;
    bool changed = !(bot && A.stop_on_bot) && np > 0;
    while(changed) {
      // The root is a common fixpoint of the table (k_pack_table checked), so in the first sweep of a subproblem only the
      // propagators that mention a halved decision variable can move anything: a few hundred records instead of all.
      const unsigned a_sbot = (sweeps & 1) ? a_sbot1 : a_sbot0;   // this sweep's bot word
      const int f = (np1 > 0 && sweeps == 0) ? pk_sweep<HAS_DIV, JOIN>(*sh1, a_T1, a_S, tid, nthr, fin, a_sbot, A.stop_on_bot, nev)
                                                    : pk_sweep<HAS_DIV, JOIN>(*sh, a_T, a_S, tid, nthr, fin, a_sbot, A.stop_on_bot, nev);
      ++sweeps;
      const int any_chg = gbar_or(bid, nthr, f & 1);   // the bot word was written where the variable was emptied
      bot |= lds_s32(a_sbot) != 0;   // complete: every thread of the group has finished the sweep that wrote it
      changed = any_chg && !(bot && A.stop_on_bot) && !(A.max_sweeps && sweeps >= A.max_sweeps);
    }
    nev = __reduce_add_sync(0xffffffffu, nev);
    if((tid & 31) == 0 && nev) atomicAdd(s_nev, nev);
    int all_ent = 0;
    if(!bot) {   // entailment: the ask loop of is_extractable over the propagators that were not entailed on the root
      int ok = 1;
      for(int i = tid; i < np && ok; i += nthr) {
        const uint2 rc = *reinterpret_cast<const uint2*>(tb + 8 * (size_t)i);
        const int2 a = lds_itv(a_S + (rc.x & 0xffffu)), bb = lds_itv(a_S + (rc.x >> 16)), c = lds_itv(a_S + (rc.y & 0xffffu));
        ok = ask_regs((int)(rc.y >> 16), Itv(a.x, a.y), Itv(bb.x, bb.y), Itv(c.x, c.y));
      }
      all_ent = gbar_and(bid, nthr, ok);
      // overflow hazard (lpc.h): a finite bound next to the int32 limits in a store that is handed back (bounds only
      // tighten: a root without a bound in that band cannot grow one)
      if(!(EPS && rflags >= 0 && !(rflags & 4)))
        for(int v = tid; v < A.nvars; v += nthr) { const int2 d = S[v]; hz |= near_inf_lo(d.x) | near_inf_hi(d.y); }
    }
    if(hz) atomicOr(&A.ctl->hazard, 1);
    fence_async_smem();
    gbar_sync(bid, nthr);
    const int nxt = *s_next;
    if(tid == 0) {
      // write-back: a failed store has no specified contents, so only the others travel
      if(!bot) {
        char* dst = nullptr;
        if(EPS) {
          const int slot = atomicAdd(&A.ctl->n_surv, 1);
          if(slot < A.surv_cap) { dst = (char*)(A.surv + (size_t)slot * store_stride); A.surv_idx[slot] = cur; }
        }
        else dst = (char*)(A.stores + cur * store_stride);
        if(dst) {
          for(int o = 0; o < sbytes; o += 32768) bulk_s2g(dst + o, (char*)S + o, min(32768, sbytes - o));
          bulk_commit();
        }
      }
      A.flags[cur] = (uint8_t)((bot ? 1 : 0) | (all_ent ? 2 : 0));
      if(A.sweeps_out) A.sweeps_out[cur] = sweeps;
      const int olb = A.objective_var >= 0 ? S[A.objective_var].x : LPC_INF;
      if(A.obj_out) A.obj_out[cur] = olb;
      if(bot) ++ga->nbot; else if(all_ent) ++ga->sol; else ++ga->unk;
      if(!bot && A.objective_var >= 0) ga->best = min(ga->best, olb);
      ga->sweeps += sweeps;
      ga->ded += *s_nev;   // complete after the barrier above: every warp has added its share
      ga->maxsw = max(ga->maxsw, sweeps);
      if(nxt >= 0) {   // the next store goes into the same slot once the write-back has read it
        bulk_wait_read0();
        mbar_expect_tx(&bars[grp], (unsigned)sbytes);
        bulk_g2s_chunked((char*)S, EPS ? (const char*)A.root : (const char*)(A.stores + nxt * store_stride), sbytes, &bars[grp]);
      }
    }
    cur = nxt;
    gbar_sync(bid, nthr);   // everyone of this group has read s_next before its thread 0 overwrites it
  }
  if(tid == 0) {
    bulk_wait0();
    if(ga->sol) atomicAdd((unsigned long long*)&A.ctl->red[0], (unsigned long long)ga->sol);
    if(ga->nbot) atomicAdd((unsigned long long*)&A.ctl->red[1], (unsigned long long)ga->nbot);
    if(ga->unk) atomicAdd((unsigned long long*)&A.ctl->red[2], (unsigned long long)ga->unk);
    atomicMin(&A.ctl->red[3], (long long)ga->best);
    atomicAdd((unsigned long long*)&A.ctl->sweeps_total, (unsigned long long)ga->sweeps);
    atomicAdd((unsigned long long*)&A.ctl->deductions, (unsigned long long)ga->ded);
    atomicMax(&A.ctl->max_sweeps_seen, ga->maxsw);
    __threadfence();
    // The group that finishes last in its block reports the block, and the block that finishes last publishes the
    // all-reduce payload (BatchCtl::payload). No block-wide barrier here: the other groups are still on their named ones.
    const bool last_group = atomicAdd(s_groups_done, 1) == G - 1;
    const int done = last_group ? atomicAdd(&A.ctl->done_blocks, 1) : -1;
    if(done == (int)gridDim.x - 1) {
      __threadfence();
      unsigned long long* red = reinterpret_cast<unsigned long long*>(A.ctl->red);   // read where the atomics landed (L2)
      long long rec[4];
      for(int i = 0; i < 4; ++i) rec[i] = (long long)atomicAdd(&red[i], 0ull);
      for(int i = 0; i < 3; ++i) A.ctl->payload[i] = rec[i];
      A.ctl->payload[3 + A.ctl->rank] = rec[3];
      if(A.peers) {   // the record goes into every rank's inbox (PeerSlot above)
        for(int r = 0; r < A.world; ++r) {
          PeerSlot* dst = A.peers[r] + (A.epoch & 1) * LPC_MAX_RANKS + A.my_rank;
          for(int i = 0; i < 4; ++i) dst->rec[i] = rec[i];
        }
        __threadfence_system();
        for(int r = 0; r < A.world; ++r) st_release_sys(&(A.peers[r] + (A.epoch & 1) * LPC_MAX_RANKS + A.my_rank)->epoch, A.epoch);
      }
    }
  }
}

} // namespace lpc

using namespace lpc;

// ---- launch plan shared by the resident and the EPS entry points -------------------------------------------------------
struct GroupPlan { int g = 0; size_t smem = 0; int sms = 0; size_t ptab_bytes = 0; int cap1 = 0; };   // cap1: records of the first-sweep table

// The join variant (pk_rule) goes with the schedule. Dense sweeps (LPC_MODE_SWEEP) join with guarded atomics (JOIN = 1);
// the default mode, which evaluates only the propagators not entailed on the root and so moves a bound in a larger share
// of its evaluations, joins with plain predicated stores inside the moved-region (JOIN = 4): 2-4 % more sweeps, each of
// them cheaper. Measured on config 4, dense / AUTO, ms (profiles/r02_ab_join.txt): JOIN 0 = 4.63 / 3.24, 1 = 4.36 / 3.20,
// 2 = 5.12 / 3.35, 3 = 5.37 / 3.18, 4 = 4.45 / 3.06. LPC_JOIN forces one variant for A/B runs (0, 2, 3 are built for the
// 8-group kernel without divisions only).
static int join_of_mode(int mode) {
  static int forced = -2;
  if(forced == -2) { const char* e = getenv("LPC_JOIN"); forced = e ? atoi(e) : -1; }
  return forced >= 0 ? forced : (mode == LPC_MODE_SWEEP ? 1 : 4);
}

template <bool EPS>
static const void* group_kernel(bool has_div, int g, int join) {
  if(!has_div && g == 8 && join == 0) return (const void*)k_pir_group<false, 8, EPS, 0>;
  if(!has_div && g == 8 && join == 2) return (const void*)k_pir_group<false, 8, EPS, 2>;
  if(!has_div && g == 8 && join == 3) return (const void*)k_pir_group<false, 8, EPS, 3>;
  if(g == 8 && join == 4) return has_div ? (const void*)k_pir_group<true, 8, EPS, 4> : (const void*)k_pir_group<false, 8, EPS, 4>;
  switch(g) {
    case 8: return has_div ? (const void*)k_pir_group<true, 8, EPS> : (const void*)k_pir_group<false, 8, EPS>;
    case 4: return has_div ? (const void*)k_pir_group<true, 4, EPS> : (const void*)k_pir_group<false, 4, EPS>;
    case 2: return has_div ? (const void*)k_pir_group<true, 2, EPS> : (const void*)k_pir_group<false, 2, EPS>;
    default: return nullptr;
  }
}

static size_t packed_capacity(const lpc_table* t) {   // records + one padding duplicate per run, rounded to a pair
  return (size_t)(t->dev.n + LPC_PK_MAXRUN + 2) / 2 * 2;
}

// The most groups whose store slots fit next to the (full) packed table. plan->g == 0: does not qualify.
template <bool EPS>
static int group_plan(const lpc_table* t, int nvars, int sbytes, GroupPlan* plan) {
  plan->g = 0;
  if(nvars > 8191 || t->dev.n >= (1 << 24)) return LPC_OK;
  int dev = 0, sms = 0, optin = 0;
  LPC_CUDA(cudaGetDevice(&dev));
  LPC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  LPC_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  plan->sms = sms;
  plan->ptab_bytes = packed_capacity(t) * 8;
  const char* e = getenv("LPC_BATCH_DUAL");
  const int want = e ? atoi(e) : -1;
  static const int kG[3] = {8, 4, 2};
  for(int c = 0; c < 3; ++c) {
    const int g = want > 0 ? want : kG[c];
    const void* k = group_kernel<EPS>(t->has_div, g, join_of_mode(LPC_MODE_SWEEP));
    const void* k2 = group_kernel<EPS>(t->has_div, g, join_of_mode(LPC_MODE_AUTO));
    const size_t need = 1024 + (size_t)g * sbytes + plan->ptab_bytes;
    if(k && need <= (size_t)optin) {
      LPC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
      if(k2 != k) LPC_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
      int per_sm = 0;
      LPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, 1024, need));
      if(per_sm >= 1) {
        plan->g = g; plan->smem = need;
        {   // what is left of the SM's shared memory takes the first-sweep table (up to 2,048 records)
          const size_t room = ((size_t)optin - need) / 16 * 16;
          const int cap1 = (int)std::min<size_t>(room / 8, 2048);
          if(cap1 >= 64) {
            plan->cap1 = cap1; plan->smem = need + (size_t)cap1 * 8;
            LPC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->smem));
            if(k2 != k) LPC_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->smem));
            LPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, 1024, plan->smem));
            if(per_sm < 1) { plan->cap1 = 0; plan->smem = need; }
          }
        }
        return LPC_OK;
      }
    }
    if(want > 0) break;
  }
  return LPC_OK;
}

static void ctl_init(BatchCtl* h, int next_store, int rank, int world) {
  memset(h, 0, sizeof(BatchCtl));
  h->red[3] = LPC_INF;
  h->next_store = next_store;
  h->rank = rank; h->world = world;
}

// ---- resident store images: the dense path of lpc_batch_fixpoint --------------------------------------------------------
int lpc_group_launch_resident(lpc_batch* b, const lpc_fixpoint_opts* o, int32_t objective_var, int first, int count,
                              BatchCtl* d_ctl, BatchCtl* h_init, cudaStream_t st, int* used) {
  *used = 0;
  const lpc_table* t = b->table;
  if(b->grp_g < 0) {
    GroupPlan p;
    int rc = group_plan<false>(t, b->nvars, b->sbytes, &p);
    if(rc) return rc;
    const char* ev = getenv("LPC_BATCH_V2");
    if(ev && !atoi(ev)) p.g = 0;
    b->grp_g = p.g; b->grp_smem = p.smem; b->grp_ptab_bytes = p.ptab_bytes; b->grp_cap1 = p.cap1;
    if(p.g) {
      b->grp_grid = std::max(1, std::min((b->n_stores + p.g - 1) / p.g, p.sms));
      LPC_CUDA(cudaMalloc(&b->d_ptab, p.ptab_bytes + (size_t)p.cap1 * 8));
      LPC_CUDA(cudaMalloc(&b->d_phdr, 2 * sizeof(PackedHdr)));
    }
  }
  // small batches keep the one-store-per-block kernel (more threads per store finish a single store sooner)
  if(b->grp_g == 0 || t->dev.n_pad < 2048 || b->n_stores < 8 * b->table->sm_count || count <= 0) return LPC_OK;
  // LPC_MODE_AUTO on a batch whose stores are tightenings of a known root: propagators entailed on the root are dropped
  const int2* root = (o->mode == LPC_MODE_AUTO && b->root_valid) ? b->d_root : nullptr;
  // ... and while every image still is the root outside the decision variables of the split (no fixpoint, no write since),
  // the first sweep may run on the records of those variables only (k_pack_table checks that the root is a fixpoint)
  const bool fresh = root && b->split_fresh && b->n_split_vars > 0 && b->grp_cap1 > 0;
  uint2* ptab1 = fresh ? (uint2*)((char*)b->d_ptab + b->grp_ptab_bytes) : nullptr;
  b->split_fresh = false;   // whatever mode this call runs in, it moves the images off the root
  k_pack_table<<<1, PK_T, 0, st>>>(t->dev, t->opsegs, root, (uint2*)b->d_ptab, (PackedHdr*)b->d_phdr, b->d_split_vars, b->n_split_vars,
                                   ptab1, b->grp_cap1, nullptr);
  g_launches++;
  LPC_CUDA(cudaGetLastError());
  const int g = b->grp_g;
  const int grid = std::max(1, std::min(b->grp_grid, (count + g - 1) / g));
  ctl_init(h_init, g * grid, b->rank, b->world);
  LPC_CUDA(cudaMemcpyAsync(d_ctl, h_init, sizeof(BatchCtl), cudaMemcpyHostToDevice, st));
  GroupArgs A{};
  A.ptab = (const uint2*)b->d_ptab; A.hdr = (const PackedHdr*)b->d_phdr; A.ptab1 = ptab1;
  A.stores = b->d + (size_t)first * b->nvars;
  A.n_stores = count; A.nvars = b->nvars; A.sbytes = b->sbytes;
  A.flags = b->d_flags + first; A.sweeps_out = b->d_sweeps + first; A.obj_out = b->d_obj ? b->d_obj + first : nullptr;
  A.ctl = d_ctl; A.objective_var = objective_var; A.max_sweeps = o->max_sweeps; A.stop_on_bot = o->stop_on_bot;
  void* args[] = {&A};
  LPC_CUDA(cudaLaunchKernel(group_kernel<false>(t->has_div, g, join_of_mode(o->mode)), dim3(grid), dim3(1024), args, b->grp_smem, st));
  g_launches++;
  *used = 1;
  return LPC_OK;
}

// ---- the EPS-native handle ----------------------------------------------------------------------------------------------
struct lpc_eps {
  const lpc_table* table = nullptr;
  int device = 0, max_n = 0, nvars = 0, sbytes = 0, surv_cap = 0;
  int2* d_root = nullptr; int* d_dvars = nullptr; long long* d_ids = nullptr;
  uint8_t* d_flags = nullptr; int* d_sweeps = nullptr; int* d_obj = nullptr;
  int2* d_surv = nullptr; int* d_surv_idx = nullptr;
  BatchCtl* d_ctl = nullptr; BatchCtl* h_ctl = nullptr; BatchCtl* h_init = nullptr;
  void* d_ptab = nullptr; PackedHdr* d_phdr = nullptr; PackedHdr* h_phdr = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStream_t last_stream = nullptr;
  bool pending = false, have_ids = false, uploaded = false;
  // The packed tables (k_pack_table) are a function of the table, the root, the decision list and the mode: they are kept
  // from one run to the next until lpc_eps_upload brings a new problem (-1: none packed yet)
  int packed_mode = -1;
  long long first_id = 0;
  int n = 0, ndec = 0;
  GroupPlan plan;
  int rank = 0, world = 1;
  long long table_gen = 0;
  // lpc_eps_solve_host with pinned, device-accessible output buffers: the kernel writes the surviving stores straight
  // into the caller's host memory (zero copy), so that their trip over the link overlaps the fixpoints
  int2* direct_surv = nullptr; int* direct_idx = nullptr; int direct_cap = 0;
  // record exchange over peer memory (PeerSlot): this rank's inbox, the peers' inboxes as mapped here, the step counter
  PeerSlot* d_inbox = nullptr;
  PeerSlot** d_peer_ptrs = nullptr;
  void* peer_mapped[LPC_MAX_RANKS] = {nullptr};
  bool peers_ok = false;
  long long epoch = 0;
};

extern "C" {

int lpc_eps_destroy(lpc_eps* e) {
  if(!e) return LPC_OK;
  cudaFree(e->d_root); cudaFree(e->d_dvars); cudaFree(e->d_ids); cudaFree(e->d_flags); cudaFree(e->d_sweeps); cudaFree(e->d_obj);
  cudaFree(e->d_surv); cudaFree(e->d_surv_idx); cudaFree(e->d_ctl); cudaFree(e->d_ptab); cudaFree(e->d_phdr);
  for(int r = 0; r < LPC_MAX_RANKS; ++r) if(e->peer_mapped[r]) cudaIpcCloseMemHandle(e->peer_mapped[r]);
  cudaFree(e->d_inbox); cudaFree(e->d_peer_ptrs);
  if(e->h_ctl) cudaFreeHost(e->h_ctl);
  if(e->h_init) cudaFreeHost(e->h_init);
  if(e->h_phdr) cudaFreeHost(e->h_phdr);
  if(e->ev0) cudaEventDestroy(e->ev0);
  if(e->ev1) cudaEventDestroy(e->ev1);
  delete e;
  return LPC_OK;
}

int lpc_eps_create(const lpc_table* t, int32_t max_subproblems, int32_t survivor_cap, lpc_eps** out) {
  LPC_REQUIRE(t && out, "null argument");
  LPC_REQUIRE(max_subproblems >= 0 && survivor_cap >= 0, "bad size");
  const int nvars = t->dev.nvars;
  if((nvars * 8) % 16 != 0) {
    set_error("lpc_eps_create: stores need an even number of variables (got %d); pad the model with one unused variable", nvars);
    return LPC_ERR_UNSUPPORTED;
  }
  LPC_REQUIRE(t->finalized, "lpc_table_finalize has not been called since the last change of the table");
  lpc_eps* e = new lpc_eps();
  e->table_gen = t->generation;
  e->table = t; e->max_n = max_subproblems; e->nvars = nvars; e->sbytes = nvars * 8; e->surv_cap = survivor_cap;
  int rc = LPC_OK;
  auto fail = [&](int code) { lpc_eps_destroy(e); return code; };
#define LPC_TRY(call) do { cudaError_t e__ = (call); if(e__ != cudaSuccess) return fail(lpc::cuda_fail(e__, #call, __FILE__, __LINE__)); } while(0)
  LPC_TRY(cudaGetDevice(&e->device));
  rc = group_plan<true>(t, nvars, e->sbytes, &e->plan);
  if(rc) return fail(rc);
  if(e->plan.g == 0) {
    set_error("lpc_eps_create: the model (%d variables, %lld propagators) does not fit the shared memory of an SM; use "
              "lpc_batch_* with resident stores", nvars, (long long)t->dev.n);
    return fail(LPC_ERR_UNSUPPORTED);
  }
  const size_t n1 = (size_t)std::max(max_subproblems, 16);
  LPC_TRY(cudaMalloc((void**)&e->d_root, std::max<size_t>(e->sbytes, 16)));
  LPC_TRY(cudaMalloc((void**)&e->d_dvars, 64 * sizeof(int)));
  LPC_TRY(cudaMalloc((void**)&e->d_ids, n1 * 8));
  LPC_TRY(cudaMalloc((void**)&e->d_flags, n1));
  LPC_TRY(cudaMalloc((void**)&e->d_sweeps, n1 * 4));
  LPC_TRY(cudaMalloc((void**)&e->d_obj, n1 * 4));
  LPC_TRY(cudaMalloc((void**)&e->d_surv, std::max<size_t>((size_t)survivor_cap * e->sbytes, 16)));
  LPC_TRY(cudaMalloc((void**)&e->d_surv_idx, std::max<size_t>((size_t)survivor_cap * 4, 16)));
  LPC_TRY(cudaMalloc((void**)&e->d_ctl, sizeof(BatchCtl)));
  LPC_TRY(cudaMalloc(&e->d_ptab, e->plan.ptab_bytes + (size_t)e->plan.cap1 * 8));
  LPC_TRY(cudaMalloc((void**)&e->d_phdr, 2 * sizeof(PackedHdr)));
  LPC_TRY(cudaHostAlloc((void**)&e->h_ctl, sizeof(BatchCtl), cudaHostAllocDefault));
  LPC_TRY(cudaHostAlloc((void**)&e->h_init, sizeof(BatchCtl), cudaHostAllocDefault));
  LPC_TRY(cudaHostAlloc((void**)&e->h_phdr, 2 * sizeof(PackedHdr), cudaHostAllocDefault));
  LPC_TRY(cudaEventCreate(&e->ev0));
  LPC_TRY(cudaEventCreate(&e->ev1));
#undef LPC_TRY
  memset(e->h_ctl, 0, sizeof(BatchCtl));
  memset(e->h_phdr, 0, 2 * sizeof(PackedHdr));
  *out = e;
  return LPC_OK;
}

int lpc_eps_set_rank(lpc_eps* e, int32_t rank, int32_t world) {
  LPC_REQUIRE(e && world >= 1 && world <= LPC_MAX_RANKS && rank >= 0 && rank < world, "bad rank / world");
  e->rank = rank; e->world = world;
  return LPC_OK;
}

void* lpc_eps_payload_device_ptr(lpc_eps* e, int32_t* n_int64) {
  if(!e) return nullptr;
  if(n_int64) *n_int64 = 3 + e->world;
  return (void*)e->d_ctl->payload;
}

static int eps_check_device(const lpc_eps* e) {
  int dev = -1;
  LPC_CUDA(cudaGetDevice(&dev));
  LPC_REQUIRE(dev == e->device, "the handle lives on another CUDA device than the current one");
  return LPC_OK;
}

static int eps_upload(lpc_eps* e, const int32_t* root, const int32_t* dvars, int32_t ndec, const int64_t* ids, int64_t first_id,
                      int32_t n, cudaStream_t st) {
  LPC_REQUIRE(e && root && (ndec == 0 || dvars), "null argument");
  LPC_REQUIRE(!e->pending, "a call is still in flight on this handle (collect it first)");
  LPC_REQUIRE(n >= 0 && n <= e->max_n, "more subproblems than the handle was created for");
  LPC_REQUIRE(ndec >= 0 && ndec < 63, "bad n_decisions");
  for(int j = 0; j < ndec; ++j) {
    LPC_REQUIRE(dvars[j] >= 0 && dvars[j] < e->nvars, "decision variable out of range");
    for(int k = 0; k < j; ++k) LPC_REQUIRE(dvars[k] != dvars[j], "decision variables must be distinct");
  }
  int rc = eps_check_device(e);
  if(rc) return rc;
  if(e->sbytes) LPC_CUDA(cudaMemcpyAsync(e->d_root, root, e->sbytes, cudaMemcpyHostToDevice, st));
  if(ndec) LPC_CUDA(cudaMemcpyAsync(e->d_dvars, dvars, (size_t)ndec * 4, cudaMemcpyHostToDevice, st));
  if(ids && n) LPC_CUDA(cudaMemcpyAsync(e->d_ids, ids, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  e->have_ids = ids != nullptr; e->first_id = first_id; e->n = n; e->ndec = ndec; e->uploaded = true;
  e->packed_mode = -1;
  return LPC_OK;
}

int lpc_eps_upload(lpc_eps* e, const int32_t* root_lbub, const int32_t* decision_vars, int32_t n_decisions,
                   const int64_t* ids, int64_t first_id, int32_t n) {
  int rc = eps_upload(e, root_lbub, decision_vars, n_decisions, ids, first_id, n, nullptr);
  if(rc) return rc;
  LPC_CUDA(cudaStreamSynchronize(nullptr));
  return LPC_OK;
}

int lpc_eps_run_async(lpc_eps* e, const lpc_fixpoint_opts* o, int32_t objective_var) {
  LPC_REQUIRE(e != nullptr && e->uploaded, "no problem uploaded");
  LPC_REQUIRE(!e->pending, "a call is still in flight on this handle (collect it first)");
  LPC_REQUIRE(objective_var < e->nvars, "objective variable out of range");
  LPC_REQUIRE(e->table->generation == e->table_gen, "the table changed since this handle was created: create a new one");
  int rc = eps_check_device(e);
  if(rc) return rc;
  lpc_fixpoint_opts def;
  if(!o) { lpc_fixpoint_default_opts(&def); o = &def; }
  cudaStream_t st = (cudaStream_t)o->stream;
  const lpc_table* t = e->table;
  const int g = e->plan.g;
  const int grid = std::max(1, std::min((e->n + g - 1) / g, e->plan.sms));
  LPC_CUDA(cudaEventRecord(e->ev0, st));
  ctl_init(e->h_init, g * grid, e->rank, e->world);
  LPC_CUDA(cudaMemcpyAsync(e->d_ctl, e->h_init, sizeof(BatchCtl), cudaMemcpyHostToDevice, st));
  // LPC_MODE_SWEEP keeps every propagator (the reference's work unit: each sweep evaluates all of them); the default mode
  // drops the ones entailed on the root
  const int2* elim_root = o->mode == LPC_MODE_SWEEP ? nullptr : e->d_root;
  uint2* ptab1 = (elim_root && e->plan.cap1 > 0 && e->ndec > 0) ? (uint2*)((char*)e->d_ptab + e->plan.ptab_bytes) : nullptr;
  const int pack_mode = o->mode == LPC_MODE_SWEEP ? 1 : 0;
  if(e->packed_mode != pack_mode) {
    k_pack_table<<<1, PK_T, 0, st>>>(t->dev, t->opsegs, elim_root, (uint2*)e->d_ptab, e->d_phdr, e->d_dvars, e->ndec, ptab1, e->plan.cap1, e->d_root);
    g_launches++;
    e->packed_mode = pack_mode;
  }
  LPC_CUDA(cudaGetLastError());
  if(e->n > 0) {
    GroupArgs A{};
    A.ptab = (const uint2*)e->d_ptab; A.hdr = e->d_phdr; A.ptab1 = ptab1;
    A.root = e->d_root; A.dvars = e->d_dvars; A.ndec = e->ndec; A.ids = e->have_ids ? e->d_ids : nullptr; A.first_id = e->first_id;
    A.n_stores = e->n; A.nvars = e->nvars; A.sbytes = e->sbytes;
    A.flags = e->d_flags; A.sweeps_out = e->d_sweeps; A.obj_out = e->d_obj;
    A.surv = e->d_surv; A.surv_idx = e->d_surv_idx; A.surv_cap = e->surv_cap;
    if(e->direct_surv) { A.surv = e->direct_surv; A.surv_idx = e->direct_idx; A.surv_cap = e->direct_cap; }
    A.ctl = e->d_ctl; A.objective_var = objective_var; A.max_sweeps = o->max_sweeps; A.stop_on_bot = o->stop_on_bot;
    ++e->epoch;
    if(e->peers_ok) { A.peers = e->d_peer_ptrs; A.my_rank = e->rank; A.world = e->world; A.epoch = e->epoch; }
    void* args[] = {&A};
    LPC_CUDA(cudaLaunchKernel(group_kernel<true>(t->has_div, g, join_of_mode(o->mode)), dim3(grid), dim3(1024), args, e->plan.smem, st));
    g_launches++;
  }
  LPC_CUDA(cudaEventRecord(e->ev1, st));
  if(e->peers_ok && e->n > 0) {   // wait for the peers' records and fold them into the payload (PeerSlot)
    k_eps_fold<<<1, 1, 0, st>>>(e->d_inbox, e->world, e->epoch, e->d_ctl);
    g_launches++;
    LPC_CUDA(cudaGetLastError());
  }
  LPC_CUDA(cudaMemcpyAsync(e->h_ctl, e->d_ctl, sizeof(BatchCtl), cudaMemcpyDeviceToHost, st));
  LPC_CUDA(cudaMemcpyAsync(e->h_phdr, e->d_phdr, 2 * sizeof(PackedHdr), cudaMemcpyDeviceToHost, st));
  e->last_stream = st;
  e->pending = true;
  return LPC_OK;
}

int lpc_eps_collect(lpc_eps* e, lpc_eps_result* r) {
  LPC_REQUIRE(e != nullptr, "null handle");
  LPC_REQUIRE(e->pending, "no call in flight on this handle");
  LPC_CUDA(cudaStreamSynchronize(e->last_stream));
  e->pending = false;
  if(r) {
    memset(r, 0, sizeof(*r));
    const BatchCtl& h = *e->h_ctl;
    r->n_solution = h.red[0]; r->n_bot = h.red[1]; r->n_unknown = h.red[2];
    r->best_bound = (int32_t)h.red[3];
    r->max_sweeps_seen = h.max_sweeps_seen;
    r->sweeps_total = h.sweeps_total; r->deductions = h.deductions;
    r->n_survivors = h.n_surv;
    r->n_live_records = e->h_phdr[0].n_live;
    r->n_first_sweep_records = e->h_phdr[1].np;
    r->overflow_hazard = h.hazard;
    float ms = 0;
    LPC_CUDA(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    r->device_ms = ms;
  }
  return LPC_OK;
}

int lpc_eps_download(lpc_eps* e, uint8_t* flags, int32_t* survivors_lbub, int32_t* survivor_index, int32_t max_survivors,
                     int32_t* n_written) {
  LPC_REQUIRE(e != nullptr && !e->pending, "collect the call first");
  LPC_REQUIRE(max_survivors >= 0, "bad max_survivors");
  cudaStream_t st = e->last_stream;
  const int ns = std::min(std::min(e->h_ctl->n_surv, e->surv_cap), max_survivors);
  if(flags && e->n) LPC_CUDA(cudaMemcpyAsync(flags, e->d_flags, e->n, cudaMemcpyDeviceToHost, st));
  if(survivors_lbub && ns) LPC_CUDA(cudaMemcpyAsync(survivors_lbub, e->d_surv, (size_t)ns * e->sbytes, cudaMemcpyDeviceToHost, st));
  if(survivor_index && ns) LPC_CUDA(cudaMemcpyAsync(survivor_index, e->d_surv_idx, (size_t)ns * 4, cudaMemcpyDeviceToHost, st));
  LPC_CUDA(cudaStreamSynchronize(st));
  if(n_written) *n_written = ns;
  return LPC_OK;
}

// The device address of a host buffer the GPU can write directly (pinned and mapped), or null.
static void* device_view_of_pinned(const void* p) {
  if(!p) return nullptr;
  cudaPointerAttributes at;
  if(cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if(at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
  return at.devicePointer;
}

int lpc_eps_solve_host(lpc_eps* e, const int32_t* root_lbub, const int32_t* decision_vars, int32_t n_decisions,
                       const int64_t* ids, int64_t first_id, int32_t n, const lpc_fixpoint_opts* o, int32_t objective_var,
                       uint8_t* flags, int32_t* survivors_lbub, int32_t* survivor_index, int32_t max_survivors,
                       int32_t* n_written, lpc_eps_result* r) {
  LPC_REQUIRE(max_survivors >= 0, "bad max_survivors");
  cudaStream_t st = o ? (cudaStream_t)o->stream : nullptr;
  int rc = eps_upload(e, root_lbub, decision_vars, n_decisions, ids, first_id, n, st);
  if(rc) return rc;
  // Pinned output buffers: the kernel stores the survivors into them directly (one bulk store each, over the link while
  // the other stores are still being propagated) instead of compacting them in device memory for a copy afterwards.
  const char* ez = getenv("LPC_EPS_ZEROCOPY");
  void* dv_surv = (!ez || atoi(ez)) ? device_view_of_pinned(survivors_lbub) : nullptr;
  void* dv_idx = dv_surv ? device_view_of_pinned(survivor_index) : nullptr;
  const bool direct = dv_surv && dv_idx && max_survivors > 0;
  if(direct) { e->direct_surv = (int2*)dv_surv; e->direct_idx = (int*)dv_idx; e->direct_cap = max_survivors; }
  rc = lpc_eps_run_async(e, o, objective_var);
  e->direct_surv = nullptr; e->direct_idx = nullptr; e->direct_cap = 0;
  if(rc) return rc;
  // the flags do not depend on the survivor count: queue their copy behind the kernel, before the first synchronisation
  if(flags && n) LPC_CUDA(cudaMemcpyAsync(flags, e->d_flags, n, cudaMemcpyDeviceToHost, st));
  if((rc = lpc_eps_collect(e, r))) return rc;
  if(direct) {
    if(n_written) *n_written = std::min(e->h_ctl->n_surv, max_survivors);
    return LPC_OK;
  }
  return lpc_eps_download(e, nullptr, survivors_lbub, survivor_index, max_survivors, n_written);
}

int lpc_eps_peer_export(lpc_eps* e, void* handle_out) {
  LPC_REQUIRE(e && handle_out, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) <= LPC_PEER_HANDLE_BYTES, "IPC handle size");
  int rc = eps_check_device(e);
  if(rc) return rc;
  if(!e->d_inbox) {
    LPC_CUDA(cudaMalloc((void**)&e->d_inbox, 2 * LPC_MAX_RANKS * sizeof(PeerSlot)));
    LPC_CUDA(cudaMemset(e->d_inbox, 0, 2 * LPC_MAX_RANKS * sizeof(PeerSlot)));
    LPC_CUDA(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  LPC_CUDA(cudaIpcGetMemHandle(&h, e->d_inbox));
  memset(handle_out, 0, LPC_PEER_HANDLE_BYTES);
  memcpy(handle_out, &h, sizeof(h));
  return LPC_OK;
}

int lpc_eps_peer_connect(lpc_eps* e, int32_t rank, int32_t world, const void* handles) {
  LPC_REQUIRE(e && handles, "null argument");
  LPC_REQUIRE(world >= 1 && world <= LPC_MAX_RANKS && rank >= 0 && rank < world, "bad rank / world");
  LPC_REQUIRE(e->d_inbox != nullptr, "call lpc_eps_peer_export first");
  LPC_REQUIRE(!e->pending, "a call is still in flight on this handle");
  int rc = eps_check_device(e);
  if(rc) return rc;
  PeerSlot* ptrs[LPC_MAX_RANKS];
  for(int r = 0; r < world; ++r) {
    if(r == rank) { ptrs[r] = e->d_inbox; continue; }
    if(e->peer_mapped[r]) { cudaIpcCloseMemHandle(e->peer_mapped[r]); e->peer_mapped[r] = nullptr; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * LPC_PEER_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    cudaError_t err = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if(err != cudaSuccess) {   // no peer access between the two devices: the caller keeps its all-reduce
      e->peers_ok = false;
      return cuda_fail(err, "cudaIpcOpenMemHandle", __FILE__, __LINE__);
    }
    e->peer_mapped[r] = p;
    ptrs[r] = (PeerSlot*)p;
  }
  if(!e->d_peer_ptrs) LPC_CUDA(cudaMalloc((void**)&e->d_peer_ptrs, LPC_MAX_RANKS * sizeof(PeerSlot*)));
  LPC_CUDA(cudaMemcpy(e->d_peer_ptrs, ptrs, world * sizeof(PeerSlot*), cudaMemcpyHostToDevice));
  e->rank = rank; e->world = world;
  e->epoch = 0;   // the ranks count their calls from here on: they must make the same calls in the same order
  e->peers_ok = true;
  return LPC_OK;
}

int lpc_eps_peer_disconnect(lpc_eps* e) {
  LPC_REQUIRE(e != nullptr && !e->pending, "null handle or a call in flight");
  for(int r = 0; r < LPC_MAX_RANKS; ++r) if(e->peer_mapped[r]) { cudaIpcCloseMemHandle(e->peer_mapped[r]); e->peer_mapped[r] = nullptr; }
  e->peers_ok = false;
  return LPC_OK;
}

int lpc_eps_sweeps(lpc_eps* e, int32_t* out) {
  LPC_REQUIRE(e && (out || e->n == 0) && !e->pending, "bad argument");
  if(e->n) LPC_CUDA(cudaMemcpy(out, e->d_sweeps, (size_t)e->n * 4, cudaMemcpyDeviceToHost));
  return LPC_OK;
}

} // extern "C"
