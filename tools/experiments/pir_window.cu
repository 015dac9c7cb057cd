// pir_window.cu — dense sweeps of the single-store fixpoint with the store staged in shared memory windows (sm_100a).
//
// Why: the dense sweep of pir_fixpoint.cu is bound by the L1 tag stage, not by bandwidth: a record's three 8-byte bound
// gathers land in ~20 different 128-byte lines per warp request and L1 looks up one line per cycle
// (profiles/r01_fixpoint_ncu.txt). Shared memory has no tag stage: a random 8-byte gather of a warp is a handful of
// bank wavefronts. The table is sorted by (op, y, x, z) (pir.hpp:343-347), so a block that owns a contiguous run of
// records sees a narrow range of y and, in models with locality, x and z close to it. Per run of records (a "chunk")
// the block copies the window [ymin - margin, ymax + margin] of the store into shared memory with coalesced loads,
// serves every gather that falls inside from there and the rest from L2 (ld.cg), evaluates the rules in registers and
// joins tightened bounds into BOTH the global store (atomicMax / atomicMin at L2) and the window copy, so the rest of
// the chunk sees them. Windows are re-read from the global store at every chunk of every sweep, hence after each grid
// barrier: the last, quiescent sweep evaluates every propagator on the final store.
// The chunk list (record range + window) is a launch plan computed once per table on the host.
// One block of 1024 threads per SM; sweeps end in the vote-carrying grid barrier of grid_barrier.cuh.
#include "lpc_internal.cuh"
#include "grid_barrier.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace lpc {

constexpr int WTPB = 1024;

struct WinChunk { int r_begin, r_end, wlo, wn; };   // records [r_begin, r_end) (multiples of 4), window [wlo, wlo + wn)

__device__ __forceinline__ int2 win_fetch(const int2* win, const int2* store, int v, int wlo, int wn) {
  const unsigned i = (unsigned)(v - wlo);
  return i < (unsigned)wn ? win[i] : __ldcg(&store[v]);
}

// Join into the global store and, when the variable is staged, into the window copy. bit0 = changed, bit1 = empty.
__device__ __forceinline__ int win_commit(int2* win, int2* store, int v, int wlo, int wn, int2 old, const Itv& nw) {
  int f = 0;
  const unsigned i = (unsigned)(v - wlo);
  if(nw.lb > old.x) { atomicMax(&store[v].x, nw.lb); if(i < (unsigned)wn) atomicMax(&win[i].x, nw.lb); f = 1; }
  if(nw.ub < old.y) { atomicMin(&store[v].y, nw.ub); if(i < (unsigned)wn) atomicMin(&win[i].y, nw.ub); f = 1; }
  if(f && nw.lb > nw.ub) f |= 2;
  return f;
}

template <bool HAS_DIV>
__global__ void __launch_bounds__(WTPB, 1) k_pir_window(TableDev t, int2* store, const WinChunk* chunks, const int* blk_off,
                                                        FixCtl* ctl, int max_sweeps, int stop_on_bot) {
  extern __shared__ int2 win[];
  __shared__ unsigned long long s_vote;
  const int tid = threadIdx.x;
  const long long gtid = blockIdx.x * (long long)WTPB + tid;
  const long long gthreads = (long long)gridDim.x * WTPB;
  int nbar = 0;
  bool bot;
  {
    int f = 0;
    for(long long i = gtid; i < t.nvars; i += gthreads) { int2 v = __ldcg(&store[i]); f |= v.x > v.y; }
    bot = grid_vote_barrier(ctl->bar, nbar++, false, f != 0, &s_vote).bot;
  }
  int sweeps = 0;
  bool any_changed = false;
  bool done = (bot && stop_on_bot) || t.n == 0;
  const int c0 = blk_off[blockIdx.x], c1 = blk_off[blockIdx.x + 1];
  while(!done) {
    int f = 0;
    for(int ci = c0; ci < c1; ++ci) {
      const WinChunk ch = chunks[ci];
      __syncthreads();                       // the previous chunk's readers are done with the window
      for(int i = tid; i < ch.wn; i += WTPB) win[i] = __ldcg(&store[ch.wlo + i]);
      __syncthreads();
      for(int r = ch.r_begin + 2 * tid; r < ch.r_end; r += 2 * WTPB) {
        const uchar2 o = *reinterpret_cast<const uchar2*>(t.op + r);
        const int2 X = *reinterpret_cast<const int2*>(t.x + r), Y = *reinterpret_cast<const int2*>(t.y + r),
                   Z = *reinterpret_cast<const int2*>(t.z + r);
        const int2 a0 = win_fetch(win, store, X.x, ch.wlo, ch.wn), b0 = win_fetch(win, store, Y.x, ch.wlo, ch.wn),
                   d0 = win_fetch(win, store, Z.x, ch.wlo, ch.wn);
        const int2 a1 = win_fetch(win, store, X.y, ch.wlo, ch.wn), b1 = win_fetch(win, store, Y.y, ch.wlo, ch.wn),
                   d1 = win_fetch(win, store, Z.y, ch.wlo, ch.wn);
#pragma unroll
        for(int k = 0; k < 2; ++k) {
          const int op = k ? o.y : o.x, xi = k ? X.y : X.x, yi = k ? Y.y : Y.x, zi = k ? Z.y : Z.x;
          const int2 a = k ? a1 : a0, b = k ? b1 : b0, c = k ? d1 : d0;
          Itv r1(a.x, a.y), r2(b.x, b.y), r3(c.x, c.y);
          deduce_regs<HAS_DIV>(op, r1, r2, r3);
          const bool slow = (r1.lb > a.x) | (r1.ub < a.y) | (r2.lb > b.x) | (r2.ub < b.y) | (r3.lb > c.x) | (r3.ub < c.y)
                          | (a.x > a.y) | (b.x > b.y) | (c.x > c.y);
          if(slow) {
            if((a.x > a.y) | (b.x > b.y) | (c.x > c.y)) f |= 2;
            f |= win_commit(win, store, xi, ch.wlo, ch.wn, a, r1);
            f |= win_commit(win, store, yi, ch.wlo, ch.wn, b, r2);
            f |= win_commit(win, store, zi, ch.wlo, ch.wn, c, r3);
          }
        }
      }
    }
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
    ctl->deductions = (unsigned long long)sweeps * (unsigned long long)t.n;
  }
}

} // namespace lpc

using namespace lpc;

// ---- launch plan ---------------------------------------------------------------------------------------------------------
// Cuts every opcode segment into one contiguous run of records per block (like the L1 kernel) and every run into chunks
// whose y-span plus margins fits the window. The margin is the 80th percentile of |x - y|, |z - y| over a sample of the
// table: models without locality get a margin of a few entries and mostly miss, which the caller detects through
// `hit_estimate` and answers by using the L1 kernel instead.
struct lpc_win_plan {
  int grid = 0;
  int cap = 0;                 // window capacity in variables
  size_t smem = 0;
  double hit_estimate = 0;     // sampled fraction of gathers served from the window
  int max_chunks = 0;          // per block
  void* d_chunks = nullptr;
  void* d_blk_off = nullptr;
};

void lpc_win_plan_free(lpc_win_plan* p) {
  if(!p) return;
  cudaFree(p->d_chunks); cudaFree(p->d_blk_off);
  delete p;
}

static int win_build_plan(lpc_table* t, lpc_win_plan** out) {
  lpc_win_plan* p = new lpc_win_plan();
  const long long n_pad = t->dev.n_pad, n = t->dev.n;
  const int nvars = std::max(1, t->dev.nvars);
  p->grid = (int)std::max<long long>(1, std::min<long long>(t->sm_count, (n_pad / 2 + WTPB - 1) / WTPB));
  const int cap_max = (int)((t->smem_optin - 1024) / 8);
  auto Y = [&](long long i) { return i < n ? t->host[i].y : 0; };
  // margin: 80th percentile of the operand distance to y on a sample
  std::vector<int> dist;
  const long long step = std::max<long long>(1, n / 65536);
  for(long long i = 0; i < n; i += step) {
    dist.push_back(std::abs(t->host[i].x - t->host[i].y));
    dist.push_back(std::abs(t->host[i].z - t->host[i].y));
  }
  int margin = 0;
  if(!dist.empty()) {
    std::sort(dist.begin(), dist.end());
    margin = dist[(size_t)(dist.size() * 0.8)] + 1;
  }
  margin = std::min(margin, cap_max / 3);
  // chunks: greedy extension by quads while span + 2 * margin fits
  std::vector<WinChunk> chunks;
  std::vector<int> blk_off(p->grid + 1, 0);
  int cap_used = 1;
  for(int b = 0; b < p->grid; ++b) {
    blk_off[b] = (int)chunks.size();
    for(int s = 0; s < t->seg_n; ++s) {
      const long long sq0 = t->seg_q[s], len = t->seg_q[s + 1] - sq0;
      const long long q0 = sq0 + len * b / p->grid, q1 = sq0 + len * (b + 1) / p->grid;
      long long q = q0;
      while(q < q1) {
        int lo = nvars, hi = -1;
        long long e = q;
        while(e < q1) {
          int l2 = lo, h2 = hi;
          for(int k = 0; k < 4; ++k) { const int y = Y(4 * e + k); l2 = std::min(l2, y); h2 = std::max(h2, y); }
          if(e > q && (long long)(h2 - l2 + 1) + 2LL * margin > cap_max) break;
          lo = l2; hi = h2; ++e;
        }
        const int span = hi - lo + 1;
        int wn = (int)std::min<long long>((long long)span + 2LL * margin, std::min(cap_max, nvars));
        int wlo = std::max(0, lo - (wn - span) / 2);
        if(wlo + wn > nvars) wlo = std::max(0, nvars - wn);
        chunks.push_back(WinChunk{(int)(4 * q), (int)(4 * e), wlo, wn});
        cap_used = std::max(cap_used, wn);
        q = e;
      }
    }
    p->max_chunks = std::max(p->max_chunks, (int)chunks.size() - blk_off[b]);
  }
  blk_off[p->grid] = (int)chunks.size();
  // sampled hit rate of the plan
  {
    long long hits = 0, tot = 0;
    const size_t cstep = std::max<size_t>(1, chunks.size() / 512);
    for(size_t c = 0; c < chunks.size(); c += cstep) {
      const WinChunk& ch = chunks[c];
      const int rstep = std::max(1, (ch.r_end - ch.r_begin) / 64);
      for(long long r = ch.r_begin; r < ch.r_end && r < n; r += rstep) {
        const int v[3] = {t->host[r].x, t->host[r].y, t->host[r].z};
        for(int k = 0; k < 3; ++k) { hits += (unsigned)(v[k] - ch.wlo) < (unsigned)ch.wn; ++tot; }
      }
    }
    p->hit_estimate = tot ? (double)hits / (double)tot : 0.0;
  }
  p->cap = cap_used;
  p->smem = (size_t)cap_used * 8;
  LPC_CUDA(cudaMalloc(&p->d_chunks, std::max<size_t>(chunks.size() * sizeof(WinChunk), 16)));
  LPC_CUDA(cudaMalloc(&p->d_blk_off, blk_off.size() * sizeof(int)));
  if(!chunks.empty()) LPC_CUDA(cudaMemcpy(p->d_chunks, chunks.data(), chunks.size() * sizeof(WinChunk), cudaMemcpyHostToDevice));
  LPC_CUDA(cudaMemcpy(p->d_blk_off, blk_off.data(), blk_off.size() * sizeof(int), cudaMemcpyHostToDevice));
  *out = p;
  return LPC_OK;
}

// Called by lpc_fixpoint_async for dense mode. Returns LPC_OK and sets *used = 1 when the window kernel was launched;
// *used = 0 means "plan not worthwhile, use the L1 kernel".
int lpc_win_fixpoint_launch(lpc_table* t, lpc_store* s, const lpc_fixpoint_opts* o, int* used) {
  *used = 0;
  // Opt-in (LPC_WINDOW=1): measured on config 2 the window kernel is SLOWER than the L1 kernel (61 vs 49 us per sweep):
  // the record loop takes the same time with shared-memory gathers as with L1 gathers, i.e. the dense sweep is bound by
  // instruction issue and dependent-load latency, not by the L1 tag stage, and the window copies come on top
  // (profiles/r01_summary.md). Kept, tested, as the experiment that settled that question.
  const char* e = getenv("LPC_WINDOW");
  if(!e || atoi(e) == 0) return LPC_OK;
  if(!t->win_plan_tried) {
    t->win_plan_tried = true;
    int rc = win_build_plan(t, &t->win_plan);
    if(rc) return rc;
    for(int d = 0; d < 2; ++d) {
      auto k = d ? k_pir_window<true> : k_pir_window<false>;
      LPC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t->win_plan->smem));
    }
  }
  lpc_win_plan* p = t->win_plan;
  double min_hit = 0.5;
  if(const char* e = getenv("LPC_WINDOW_MIN_HIT")) min_hit = atof(e);
  if(!p || p->hit_estimate < min_hit || p->max_chunks > 64) return LPC_OK;
  cudaStream_t st = (cudaStream_t)o->stream;
  LPC_CUDA(cudaEventRecord(s->ev0, st));
  LPC_CUDA(cudaMemsetAsync(s->d_ctl, 0, sizeof(FixCtl), st));
  TableDev td = t->dev;
  int2* store = s->d;
  const WinChunk* chunks = (const WinChunk*)p->d_chunks;
  const int* blk_off = (const int*)p->d_blk_off;
  FixCtl* ctl = s->d_ctl;
  int max_sweeps = o->max_sweeps, stop = o->stop_on_bot;
  void* args[] = {&td, &store, &chunks, &blk_off, &ctl, &max_sweeps, &stop};
  void* k = t->has_div ? (void*)k_pir_window<true> : (void*)k_pir_window<false>;
  LPC_CUDA(cudaLaunchCooperativeKernel(k, dim3(p->grid), dim3(WTPB), args, p->smem, st));
  g_launches++;
  LPC_CUDA(cudaEventRecord(s->ev1, st));
  LPC_CUDA(cudaMemcpyAsync(s->h_ctl, s->d_ctl, sizeof(FixCtl), cudaMemcpyDeviceToHost, st));
  s->last_stream = st;
  s->pending = true;
  *used = 1;
  return LPC_OK;
}
