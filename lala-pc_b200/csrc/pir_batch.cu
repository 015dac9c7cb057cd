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

#include <algorithm>
#include <cstring>

namespace lpc {

struct BatchCtl {
  long long red[4];            // n_solution, n_bot, n_unknown, best_bound (min)
  long long sweeps_total;
  long long deductions;
  int max_sweeps_seen;
  int next_store;              // dynamic scheduler
};

// ---- PTX helpers: mbarrier + bulk async copies (TMA, 1-D) --------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
    "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  while(!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// Issue a large global->shared copy as <= 32 KB bulk pieces (sizes are multiples of 16 by construction).
__device__ __forceinline__ void bulk_g2s_chunked(char* dst, const char* src, unsigned bytes, unsigned long long* bar) {
  for(unsigned o = 0; o < bytes; o += 32768u) bulk_g2s(dst + o, src + o, min(32768u, bytes - o), bar);
}

// Shared-state-space accessors with 32-bit addresses: the ring slot is selected at run time, so through generic
// pointers the compiler falls back to generic LD / ATOM and 64-bit address arithmetic.
__device__ __forceinline__ int2 lds_itv(unsigned addr) {
  int2 v;
  asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ int lds_s32(unsigned addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ int lds_u8(unsigned addr) {
  int v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void reds_max(unsigned addr, int v) { asm volatile("red.shared.max.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void reds_min(unsigned addr, int v) { asm volatile("red.shared.min.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// Join into the shared-memory store (only called when something tightened or an operand was empty).
__device__ __forceinline__ int commit_smem(unsigned addr, int2 old, const Itv& nw) {
  int f = 0;
  if(nw.lb > old.x) { reds_max(addr, nw.lb); f = 1; }
  if(nw.ub < old.y) { reds_min(addr + 4, nw.ub); f = 1; }
  if(f && nw.lb > nw.ub) f |= 2;
  return f;
}

// Shared memory carve-up (dynamic): [mbarriers 64 B][store ring 2 x sbytes][table: x | y | z | op]
template <bool HAS_DIV, bool TABLE_SMEM>
__global__ void k_pir_batch(TableDev t, int2* stores, int n_stores, int sbytes, uint8_t* flags, int* sweeps_out,
                            int* obj_out, BatchCtl* ctl, int objective_var, int max_sweeps, int stop_on_bot) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem);   // [0],[1]: store ring, [2]: table
  int* s_next = reinterpret_cast<int*>(smem + 32);
  volatile int* s_bot = reinterpret_cast<volatile int*>(smem + 36);
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

    // bot before the first sweep?
    if(tid == 0) *s_bot = 0;
    int f0 = 0;
    for(int v = tid; v < t.nvars; v += nthr) { int2 d = S[v]; f0 |= d.x > d.y; }
    bool bot = __syncthreads_or(f0) != 0;
    int sweeps = 0;
    bool changed = !(bot && stop_on_bot) && t.n > 0;
    while(changed) {
      int f = 0;
      for(int i = tid; i < npad; i += nthr) {
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
          f |= commit_smem(ax, a, r1) | commit_smem(ay, bb, r2) | commit_smem(az, c, r3);
        }
      }
      ++sweeps;
      if(f & 2) *s_bot = 1;
      // one barrier per sweep: it publishes the sweep's shared-memory joins, votes has_changed, and orders s_bot
      const int any_chg = __syncthreads_or(f & 1);
      bot |= *s_bot != 0;
      changed = any_chg && !(bot && stop_on_bot) && !(max_sweeps && sweeps >= max_sweeps);
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
      a_ded += (long long)sweeps * t.n;
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

// EPS decomposition: store k := base with decision j halved according to bit j of (first_id + k).
__global__ void k_batch_init_split(int2* stores, int nvars, int n_stores, const int2* base, const int* dvars, int ndec,
                                   long long first_id) {
  for(int k = blockIdx.x; k < n_stores; k += gridDim.x) {
    int2* S = stores + (size_t)k * nvars;
    for(int v = threadIdx.x; v < nvars; v += blockDim.x) S[v] = base[v];
    __syncthreads();
    const long long id = first_id + k;
    for(int j = threadIdx.x; j < ndec; j += blockDim.x) {
      const int v = dvars[j];
      const int2 d = base[v];
      const long long mid = (long long)d.x + (((long long)d.y - (long long)d.x) >> 1);
      S[v] = ((id >> j) & 1) ? make_int2((int)(mid + 1), d.y) : make_int2(d.x, (int)mid);
    }
    __syncthreads();
  }
}

} // namespace lpc

using namespace lpc;

struct lpc_batch {
  const lpc_table* table = nullptr;
  int n_stores = 0, nvars = 0;
  int2* d = nullptr;
  uint8_t* d_flags = nullptr;
  int* d_sweeps = nullptr;
  int* d_obj = nullptr;
  BatchCtl* d_ctl = nullptr;
  BatchCtl* h_ctl = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStream_t last_stream = nullptr;
  bool pending = false;
  int sbytes = 0;
  bool plan_ready = false, table_smem = false;
  size_t smem = 0;
  int threads = 0, grid = 0;
  lpc::BatchCtl* h_init = nullptr;   // pinned initial control block
};

typedef void (*batch_kernel_t)(TableDev, int2*, int, int, uint8_t*, int*, int*, BatchCtl*, int, int, int);

static batch_kernel_t pick_batch_kernel(bool has_div, bool table_smem) {
  if(has_div) return table_smem ? k_pir_batch<true, true> : k_pir_batch<true, false>;
  return table_smem ? k_pir_batch<false, true> : k_pir_batch<false, false>;
}

extern "C" {

int lpc_batch_create(const lpc_table* t, int32_t n_stores, lpc_batch** out) {
  LPC_REQUIRE(t && out, "null argument");
  LPC_REQUIRE(n_stores >= 0, "bad n_stores");
  lpc_batch* b = new lpc_batch();
  b->table = t; b->n_stores = n_stores; b->nvars = t->dev.nvars;
  // ring slots are multiples of 16 B (bulk copy granularity); stores are packed at nvars*8 B in global memory, so
  // the per-store image must itself be a multiple of 16 B: require an even number of variables or pad by one.
  b->sbytes = ((b->nvars * 8 + 15) / 16) * 16;
  if(b->sbytes != b->nvars * 8) {
    delete b;
    set_error("lpc_batch_create: batched stores need an even number of variables (got %d); pad the model with one unused variable", t->dev.nvars);
    return LPC_ERR_UNSUPPORTED;
  }
  size_t bytes = std::max<size_t>((size_t)n_stores * b->nvars * 8, 16);
  cudaError_t e = cudaMalloc((void**)&b->d, bytes);
  if(e != cudaSuccess) { delete b; return cuda_fail(e, "cudaMalloc(batch)", __FILE__, __LINE__); }
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
  *out = b;
  return LPC_OK;
}

int lpc_batch_destroy(lpc_batch* b) {
  if(!b) return LPC_OK;
  cudaFree(b->d); cudaFree(b->d_flags); cudaFree(b->d_sweeps); cudaFree(b->d_obj); cudaFree(b->d_ctl);
  if(b->h_ctl) cudaFreeHost(b->h_ctl);
  if(b->h_init) cudaFreeHost(b->h_init);
  if(b->ev0) cudaEventDestroy(b->ev0);
  if(b->ev1) cudaEventDestroy(b->ev1);
  delete b;
  return LPC_OK;
}

void* lpc_batch_device_ptr(lpc_batch* b) { return b ? (void*)b->d : nullptr; }
void* lpc_batch_reduction_device_ptr(lpc_batch* b) { return b ? (void*)b->d_ctl->red : nullptr; }

int lpc_batch_write(lpc_batch* b, int32_t first, int32_t n, const int32_t* lbub) {
  LPC_REQUIRE(b && (n == 0 || lbub), "null argument");
  LPC_REQUIRE(first >= 0 && n >= 0 && (long long)first + n <= b->n_stores, "range out of bounds");
  if(n) LPC_CUDA(cudaMemcpy(b->d + (size_t)first * b->nvars, lbub, (size_t)n * b->nvars * 8, cudaMemcpyHostToDevice));
  return LPC_OK;
}

int lpc_batch_read(const lpc_batch* b, int32_t first, int32_t n, int32_t* lbub) {
  LPC_REQUIRE(b && (n == 0 || lbub), "null argument");
  LPC_REQUIRE(first >= 0 && n >= 0 && (long long)first + n <= b->n_stores, "range out of bounds");
  if(n) LPC_CUDA(cudaMemcpy(lbub, b->d + (size_t)first * b->nvars, (size_t)n * b->nvars * 8, cudaMemcpyDeviceToHost));
  return LPC_OK;
}

int lpc_batch_init_split(lpc_batch* b, const int32_t* base_lbub, const int32_t* decision_vars, int32_t n_decisions,
                         int64_t first_id) {
  LPC_REQUIRE(b && base_lbub && (n_decisions == 0 || decision_vars), "null argument");
  LPC_REQUIRE(n_decisions >= 0 && n_decisions < 63, "bad n_decisions");
  for(int j = 0; j < n_decisions; ++j) LPC_REQUIRE(decision_vars[j] >= 0 && decision_vars[j] < b->nvars, "decision variable out of range");
  if(b->n_stores == 0) return LPC_OK;
  int2* d_base = nullptr;
  int* d_dec = nullptr;
  LPC_CUDA(cudaMalloc((void**)&d_base, std::max<size_t>((size_t)b->nvars * 8, 16)));
  LPC_CUDA(cudaMalloc((void**)&d_dec, std::max(n_decisions, 1) * sizeof(int)));
  LPC_CUDA(cudaMemcpy(d_base, base_lbub, (size_t)b->nvars * 8, cudaMemcpyHostToDevice));
  if(n_decisions) LPC_CUDA(cudaMemcpy(d_dec, decision_vars, n_decisions * sizeof(int), cudaMemcpyHostToDevice));
  k_batch_init_split<<<std::min(b->n_stores, 148 * 16), 256>>>(b->d, b->nvars, b->n_stores, d_base, d_dec, n_decisions, first_id);
  g_launches++;
  LPC_CUDA(cudaGetLastError());
  LPC_CUDA(cudaDeviceSynchronize());
  cudaFree(d_base);
  cudaFree(d_dec);
  return LPC_OK;
}

int lpc_batch_fixpoint_async(lpc_batch* b, const lpc_fixpoint_opts* o, int32_t objective_var) {
  LPC_REQUIRE(b != nullptr, "null batch");
  LPC_REQUIRE(objective_var < b->nvars, "objective variable out of range");
  lpc_fixpoint_opts def;
  if(!o) { lpc_fixpoint_default_opts(&def); o = &def; }
  const lpc_table* t = b->table;
  cudaStream_t st = (cudaStream_t)o->stream;
  // launch plan: computed once per batch handle (device attribute / occupancy queries are slow driver calls)
  if(!b->plan_ready) {
    int dev = 0, sms = 0, optin = 0;
    LPC_CUDA(cudaGetDevice(&dev));
    LPC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    LPC_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t max_smem = (size_t)optin;
    const size_t ring = 64 + 2 * (size_t)b->sbytes;
    const size_t tbl = (size_t)t->dev.n_pad * 13;
    if(ring > max_smem) {
      set_error("lpc_batch_fixpoint: a store of %d variables does not fit the shared-memory ring (%zu > %zu B); use lpc_fixpoint per store", b->nvars, ring, max_smem);
      return LPC_ERR_UNSUPPORTED;
    }
    // bulk copies need every sub-array 16-B aligned and sized: n_pad is a multiple of 16 (lpc_table_create)
    b->table_smem = ring + tbl <= max_smem;
    b->smem = b->table_smem ? ring + tbl : ring;
    batch_kernel_t kk = pick_batch_kernel(t->has_div, b->table_smem);
    LPC_CUDA(cudaFuncSetAttribute((const void*)kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem));
    b->threads = (int)std::min<long long>(1024, std::max<long long>(32, (t->dev.n_pad + 31) / 32 * 32));
    int per_sm = 0;
    LPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kk, b->threads, b->smem));
    if(per_sm < 1 && b->threads > 256) {
      b->threads = 256;
      LPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kk, b->threads, b->smem));
    }
    LPC_REQUIRE(per_sm > 0, "batch kernel does not fit on an SM");
    b->grid = std::max(1, std::min(b->n_stores, sms * per_sm));
    b->plan_ready = true;
  }
  batch_kernel_t k = pick_batch_kernel(t->has_div, b->table_smem);
  const int grid = b->grid, threads = b->threads;
  const size_t smem = b->smem;
  memset(b->h_init, 0, sizeof(BatchCtl));
  b->h_init->red[3] = LPC_INF;
  b->h_init->next_store = grid;
  LPC_CUDA(cudaEventRecord(b->ev0, st));
  LPC_CUDA(cudaMemcpyAsync(b->d_ctl, b->h_init, sizeof(BatchCtl), cudaMemcpyHostToDevice, st));
  if(b->n_stores > 0) {
    k<<<grid, threads, smem, st>>>(t->dev, b->d, b->n_stores, b->sbytes, b->d_flags, b->d_sweeps, b->d_obj, b->d_ctl,
                                  objective_var, o->max_sweeps, o->stop_on_bot);
    g_launches++;
    LPC_CUDA(cudaGetLastError());
  }
  LPC_CUDA(cudaEventRecord(b->ev1, st));
  LPC_CUDA(cudaMemcpyAsync(b->h_ctl, b->d_ctl, sizeof(BatchCtl), cudaMemcpyDeviceToHost, st));
  b->last_stream = st;
  b->pending = true;
  return LPC_OK;
}

int lpc_batch_collect(lpc_batch* b, lpc_batch_result* r) {
  LPC_REQUIRE(b != nullptr, "null batch");
  LPC_REQUIRE(b->pending, "no batch fixpoint in flight");
  LPC_CUDA(cudaStreamSynchronize(b->last_stream));
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
  cudaStream_t st = o ? (cudaStream_t)o->stream : nullptr;
  size_t bytes = (size_t)b->n_stores * b->nvars * 8;
  if(bytes) LPC_CUDA(cudaMemcpyAsync(b->d, lbub, bytes, cudaMemcpyHostToDevice, st));
  int rc = lpc_batch_fixpoint_async(b, o, objective_var);
  if(rc) return rc;
  if(bytes) LPC_CUDA(cudaMemcpyAsync(lbub, b->d, bytes, cudaMemcpyDeviceToHost, st));
  return lpc_batch_collect(b, r);
}

int lpc_batch_flags(const lpc_batch* b, uint8_t* out) {
  LPC_REQUIRE(b && (out || b->n_stores == 0), "null argument");
  if(b->n_stores) LPC_CUDA(cudaMemcpy(out, b->d_flags, b->n_stores, cudaMemcpyDeviceToHost));
  return LPC_OK;
}

} // extern "C"
