// pir_cluster.cu — the single-store fixpoint of a SMALL network inside one thread-block cluster (sm_100a).
//
// The grid-wide kernels (pir_fixpoint.cu, pir_dirty.cu) pay one grid barrier per sweep: ~3 us through L2 plus the
// dependent L2 round trips of a sweep that only has a few records per thread. A network whose store fits in shared
// memory does not need the grid: here ONE cluster of 8 CTAs x 1024 threads runs the whole fixpoint (config 1 of
// BASELINE.json, 10 k variables / 50 k propagators: 5.3 us per sweep with the grid kernel).
//   * every CTA keeps a full REPLICA of the store in its shared memory (loaded once by bulk copies) and one eighth of
//     the propagator table (13 B per record SoA, also loaded once): after the prologue a sweep touches no global memory;
//   * a record is evaluated on the local replica; a bound that tightens is joined into ALL replicas through distributed
//     shared memory (atomicMax / atomicMin on cluster.map_shared_rank addresses). The joins are lattice joins, so every
//     replica converges to the same store whatever the arrival order, and stale reads are harmless;
//   * the sweep ends with the cluster barrier (barrier.cluster arrive.release / wait.acquire, which also makes the remote
//     joins visible); each CTA has written its changed / bot vote into every CTA's vote row before it, so all CTAs read
//     the same verdict locally afterwards. Vote rows alternate between two buffers (a CTA can be at most one barrier
//     ahead);
//   * at the fixpoint CTA 0 writes its replica back with one bulk copy.
// Same contract as the grid kernels: final store == the Gauss-Seidel fixpoint on non-failed inputs, stop at the first
// sweep that observes bot; sweeps = cluster sweeps, deductions = sweeps x records.
// MEASURED SLOWER than the grid kernel on config 1 (see lpc_cluster_fixpoint_launch): opt-in, LPC_CLUSTER=1.
#include "lpc_internal.cuh"
#include "smem_tma.cuh"

#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace lpc {

constexpr int CL_CTAS = 8;       // portable maximum cluster size
constexpr int CL_TPB = 1024;

// Shared memory: [mbarrier 8 B | votes 2 x 8 ints | pad to 128][store replica][table slice: x | y | z | op]
template <bool HAS_DIV>
__global__ void __cluster_dims__(CL_CTAS, 1, 1) __launch_bounds__(CL_TPB, 1)
k_pir_cluster(TableDev t, int2* store, FixCtl* ctl, int slice, int sbytes, int max_sweeps, int stop_on_bot) {
  extern __shared__ __align__(128) unsigned char smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem);
  int* votes = reinterpret_cast<int*>(smem + 16);                       // [2][CL_CTAS]
  int2* S = reinterpret_cast<int2*>(smem + 128);
  char* tb = reinterpret_cast<char*>(smem + 128 + sbytes);
  const int lo = rank * slice, hi = min((int)t.n_pad, lo + slice), cnt = max(0, hi - lo);   // this CTA's records
  int* sx = reinterpret_cast<int*>(tb);
  int* sy = reinterpret_cast<int*>(tb + (size_t)slice * 4);
  int* sz = reinterpret_cast<int*>(tb + (size_t)slice * 8);
  uint8_t* sop = reinterpret_cast<uint8_t*>(tb + (size_t)slice * 12);

  if(tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if(tid < 2 * CL_CTAS) votes[tid] = 0;
  __syncthreads();
  if(tid == 0) {
    mbar_expect_tx(bar, (unsigned)(sbytes + cnt * 13));
    bulk_g2s_chunked((char*)S, (const char*)store, sbytes, bar);
    if(cnt) {
      bulk_g2s_chunked((char*)sx, (const char*)(t.x + lo), cnt * 4, bar);
      bulk_g2s_chunked((char*)sy, (const char*)(t.y + lo), cnt * 4, bar);
      bulk_g2s_chunked((char*)sz, (const char*)(t.z + lo), cnt * 4, bar);
      bulk_g2s_chunked((char*)sop, (const char*)(t.op + lo), cnt, bar);
    }
  }
  mbar_wait(bar, 0);

  // a store created with an empty variable is at bot before the first sweep (every replica sees the same store)
  int f0 = 0;
  for(int v = tid; v < t.nvars; v += CL_TPB) { const int2 d = S[v]; f0 |= d.x > d.y; }
  bool bot = __syncthreads_or(f0) != 0;
  cluster.sync();   // every replica is loaded before anyone joins into it
  int sweeps = 0;
  bool any_changed = false;
  bool done = (bot && stop_on_bot) || t.n == 0;
  while(!done) {
    int f = 0;
    for(int i = tid; i < cnt; i += CL_TPB) {
      const int op = sop[i], xi = sx[i], yi = sy[i], zi = sz[i];
      const int2 a = S[xi], b = S[yi], c = S[zi];
      Itv r1(a.x, a.y), r2(b.x, b.y), r3(c.x, c.y);
      deduce_regs<HAS_DIV>(op, r1, r2, r3);
      const bool slow = (r1.lb > a.x) | (r1.ub < a.y) | (r2.lb > b.x) | (r2.ub < b.y) | (r3.lb > c.x) | (r3.ub < c.y)
                      | (a.x > a.y) | (b.x > b.y) | (c.x > c.y);
      if(slow) {
        if((a.x > a.y) | (b.x > b.y) | (c.x > c.y)) f |= 2;
#pragma unroll
        for(int w = 0; w < 3; ++w) {
          const int v = w == 0 ? xi : w == 1 ? yi : zi;
          const int2 old = w == 0 ? a : w == 1 ? b : c;
          const Itv nw = w == 0 ? r1 : w == 1 ? r2 : r3;
          const bool cl = nw.lb > old.x, cu = nw.ub < old.y;
          if(cl | cu) {
            f |= nw.lb > nw.ub ? 3 : 1;
#pragma unroll
            for(int r = 0; r < CL_CTAS; ++r) {   // join into every replica (own included) through distributed shared memory
              int2* cell = cluster.map_shared_rank(&S[v], r);
              if(cl) atomicMax(&cell->x, nw.lb);
              if(cu) atomicMin(&cell->y, nw.ub);
            }
          }
        }
      }
    }
    // vote: this CTA's flags into row (sweeps & 1) of every CTA, then the cluster barrier
    const int any = __syncthreads_or(f & 1), anyb = __syncthreads_or(f & 2);
    if(tid < CL_CTAS) cluster.map_shared_rank(votes, tid)[(sweeps & 1) * CL_CTAS + rank] = (any ? 1 : 0) | (anyb ? 2 : 0);
    cluster.sync();
    int verdict = 0;
#pragma unroll
    for(int r = 0; r < CL_CTAS; ++r) verdict |= votes[(sweeps & 1) * CL_CTAS + r];
    ++sweeps;
    bot |= (verdict & 2) != 0;
    any_changed |= (verdict & 1) != 0;
    if(!(verdict & 1) || (bot && stop_on_bot) || (max_sweeps && sweeps >= max_sweeps)) done = true;
  }
  // no CTA may leave (and release its shared memory) while another can still join into it or read its votes
  cluster.sync();
  if(rank == 0) {
    fence_async_smem();
    __syncthreads();
    if(tid == 0) {
      for(int o = 0; o < sbytes; o += 32768) bulk_s2g((char*)store + o, (char*)S + o, min(32768, sbytes - o));
      bulk_commit();
      bulk_wait0();
      ctl->sweeps = sweeps;
      ctl->dense_sweeps = sweeps;
      ctl->has_changed = any_changed;
      ctl->is_bot = bot;
      ctl->deductions = (unsigned long long)sweeps * (unsigned long long)t.n;
    }
  }
}

} // namespace lpc

using namespace lpc;

// Called by lpc_fixpoint_async for LPC_MODE_SWEEP / LPC_MODE_AUTO. *used = 1 when the cluster kernel took the call.
int lpc_cluster_fixpoint_launch(lpc_table* t, lpc_store* s, const lpc_fixpoint_opts* o, int* used) {
  *used = 0;
  // Opt-in (LPC_CLUSTER=1): measured on config 1 the cluster kernel is SLOWER than the grid kernel - 19 us per sweep
  // against 4.9 (287 vs 106 us per fixpoint, although it needs 15 sweeps instead of 22): every tightening costs up to 16
  // remote atomics through distributed shared memory, and early sweeps tighten a third of the records. Kept, tested,
  // as the experiment that settled whether small networks should leave the grid (tools/c1_probe.py).
  const char* e = getenv("LPC_CLUSTER");
  if(!e || atoi(e) == 0) return LPC_OK;
  if(s->nvars != t->dev.nvars || t->dev.n < 1024) return LPC_OK;        // the replica is the whole store; tiny tables: one block is enough
  const int sbytes = (s->nvars * 8 + 15) / 16 * 16;
  const int slice = (int)(((t->dev.n_pad + CL_CTAS - 1) / CL_CTAS + 15) / 16 * 16);
  const size_t need = 128 + (size_t)sbytes + (size_t)slice * 13;
  if(need > t->smem_optin || (size_t)s->nvars * 8 != (size_t)sbytes) return LPC_OK;
  const void* k = t->has_div ? (const void*)k_pir_cluster<true> : (const void*)k_pir_cluster<false>;
  if(!t->cluster_ready) {
    LPC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    t->cluster_ready = true;
  }
  cudaStream_t st = (cudaStream_t)o->stream;
  LPC_CUDA(cudaEventRecord(s->ev0, st));
  LPC_CUDA(cudaMemsetAsync(s->d_ctl, 0, sizeof(FixCtl), st));
  TableDev td = t->dev;
  int2* store = s->d;
  FixCtl* ctl = s->d_ctl;
  int sl = slice, sb = sbytes, max_sweeps = o->max_sweeps, stop = o->stop_on_bot;
  void* args[] = {&td, &store, &ctl, &sl, &sb, &max_sweeps, &stop};
  LPC_CUDA(cudaLaunchKernel(k, dim3(CL_CTAS), dim3(CL_TPB), args, need, st));   // cluster dims are compiled in
  g_launches++;
  LPC_CUDA(cudaEventRecord(s->ev1, st));
  LPC_CUDA(cudaMemcpyAsync(s->h_ctl, s->d_ctl, sizeof(FixCtl), cudaMemcpyDeviceToHost, st));
  s->last_stream = st;
  s->pending = true;
  *used = 1;
  return LPC_OK;
}
