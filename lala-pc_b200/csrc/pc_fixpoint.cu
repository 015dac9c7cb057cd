// pc_fixpoint.cu — the PC fixpoint over flattened n-ary propagators (include/lpc_pc.h) for sm_100a.
//
// Replaces  GaussSeidelIteration{}.fixpoint(pc.num_deductions(), [&](size_t i){ return pc.deduce(i); }, has_changed)
// (tests/pc_test.cpp:91-94) where pc.deduce(i) walks a heap-allocated formula / term tree (pc.hpp:671-680). The flat
// table is streamed instead: one 16-byte header per propagator + 8 bytes per term, one thread per propagator, bounds
// gathered from the L2-resident store, joins by atomicMax / atomicMin. Same persistent cooperative skeleton as the
// PIR kernel: dense chaotic sweeps, block-voted has_changed / bot, a grid barrier per sweep.
#include "lpc_internal.cuh"
#include "pc_device.cuh"

#include <cooperative_groups.h>
#include <algorithm>
#include <cstring>

#include "../../include/lpc_pc.h"

namespace cg = cooperative_groups;

namespace lpc {

constexpr int PC_TPB = 256;

// VStore<Interval<ZLB>> in global memory: gathers + lattice joins at L2.
struct GlobalAcc {
  int2* s;
  mutable int seen_bot;
  __device__ __forceinline__ Itv load(int v) const {
    const int2 d = s[v];
    seen_bot |= d.x > d.y;
    return Itv(d.x, d.y);
  }
  __device__ __forceinline__ int embed(int v, const Itv& u) {
    const int2 old = s[v];
    int f = 0;
    if(u.lb > old.x) { atomicMax(&s[v].x, u.lb); f = 1; }
    if(u.ub < old.y) { atomicMin(&s[v].y, u.ub); f = 1; }
    if(f && max(u.lb, old.x) > min(u.ub, old.y)) f |= 2;
    return f;
  }
};

// VStore<NBitset<64>> in global memory: one uint64 per variable, joins by atomicAnd.
struct GlobalBitAcc {
  u64* s;
  mutable int seen_bot;
  __device__ __forceinline__ u64 load(int v) const {
    const u64 d = s[v];
    seen_bot |= d == 0;
    return d;
  }
  __device__ __forceinline__ int embed(int v, u64 u) {
    const u64 old = s[v];
    if(old == 0) return 2;
    const u64 nw = old & u;
    if(nw == old) return 0;
    atomicAnd(&s[v], u);
    return nw == 0 ? 3 : 1;
  }
};

template <bool BITS> struct PcAcc { typedef GlobalAcc type; };
template <> struct PcAcc<true> { typedef GlobalBitAcc type; };
template <bool BITS, class Acc> __device__ __forceinline__ int pc_step(Acc& acc, const int4 h, const int2* terms) {
  if constexpr(BITS) return pc_deduce_bits(acc, h, terms);
  else return pc_deduce(acc, h, terms);
}
__device__ __forceinline__ GlobalAcc make_acc(int2* store, GlobalAcc*) { return GlobalAcc{store, 0}; }
__device__ __forceinline__ GlobalBitAcc make_acc(int2* store, GlobalBitAcc*) { return GlobalBitAcc{reinterpret_cast<u64*>(store), 0}; }

template <bool BITS>
__global__ void __launch_bounds__(PC_TPB) k_pc_fixpoint(PcTableDev t, int2* store, FixCtl* ctl, int max_sweeps,
                                                        int stop_on_bot) {
  typedef typename PcAcc<BITS>::type Acc;
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x;
  const long long gtid = blockIdx.x * (long long)PC_TPB + tid;
  const long long gthreads = (long long)gridDim.x * PC_TPB;
  volatile int* vflags = ctl->flags;
  volatile int* vbot = &ctl->is_bot;
  {
    int f = 0;
    for(long long i = gtid; i < t.nvars; i += gthreads) { int2 v = store[i]; f |= BITS ? (v.x | v.y) == 0 : v.x > v.y; }
    if(__syncthreads_or(f) && tid == 0) atomicOr(&ctl->is_bot, 1);
  }
  grid.sync();
  int sweeps = 0;
  bool any_changed = false;
  bool bot = *vbot != 0;
  bool done = (bot && stop_on_bot) || t.n == 0;
  while(!done) {
    const int slot = sweeps % 3;
    if(blockIdx.x == 0 && tid == 0) vflags[(sweeps + 1) % 3] = 0;
    Acc acc = make_acc(store, (Acc*)nullptr);
    int f = 0;
    for(long long p = gtid; p < t.n; p += gthreads) {
      const int4 h = t.hdr[p];
      f |= pc_step<BITS>(acc, h, t.terms + h.y);
    }
    if(acc.seen_bot) f |= 2;
    if(__syncthreads_or(f & 2) && tid == 0) atomicOr(&ctl->is_bot, 1);
    if(__syncthreads_or(f & 1) && tid == 0) atomicOr(&ctl->flags[slot], 1);
    grid.sync();
    ++sweeps;
    const int c = vflags[slot];
    bot = *vbot != 0;
    any_changed |= c != 0;
    if(c == 0 || (bot && stop_on_bot) || (max_sweeps && sweeps >= max_sweeps)) done = true;
  }
  if(blockIdx.x == 0 && tid == 0) {
    ctl->sweeps = sweeps;
    ctl->dense_sweeps = sweeps;
    ctl->has_changed = any_changed;
    ctl->is_bot = bot;
    ctl->deductions = (unsigned long long)sweeps * (unsigned long long)t.n;
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

struct lpc_pc_table {
  PcTableDev dev{};
  void* d_hdr = nullptr;
  void* d_terms = nullptr;
  int sm_count = 0;
  int blocks_per_sm[2] = {0, 0};   // [interval store, bitset store]
  bool has_linear = false;          // LIN_LE / REIF_LIN_LE present: no bitset rule (see lpc_pc.h)
  lpc_store* host_store = nullptr;
};

extern "C" {

int lpc_pc_table_create(const lpc_pc_prop* props, int64_t n_props, const lpc_pc_term* terms, int64_t n_terms,
                        int32_t nvars, lpc_pc_table** out) {
  LPC_REQUIRE(out != nullptr, "null out");
  LPC_REQUIRE(n_props >= 0 && n_terms >= 0 && nvars >= 0, "negative size");
  LPC_REQUIRE((n_props == 0 || props) && (n_terms == 0 || terms), "null array");
  int cnt = 0;
  lpc_device_count(&cnt);
  if(cnt == 0) { set_error("no CUDA device: this library has no CPU path"); return LPC_ERR_NO_DEVICE; }
  std::vector<int4> hdr((size_t)std::max<int64_t>(n_props, 1));
  bool has_linear = false;
  for(int64_t i = 0; i < n_props; ++i) {
    const lpc_pc_prop& p = props[i];
    if(p.kind < LPC_PC_LIN_LE || p.kind > LPC_PC_ABS_EQ) { set_error("lpc_pc_table_create: propagator %lld has unsupported kind %d", (long long)i, p.kind); return LPC_ERR_UNSUPPORTED; }
    if(p.n_terms < 1 || p.n_terms >= (1 << 23) || p.first_term < 0 || (int64_t)p.first_term + p.n_terms > n_terms) { set_error("lpc_pc_table_create: propagator %lld has a bad term range", (long long)i); return LPC_ERR_INVALID; }
    const bool two = p.kind == LPC_PC_EQ || p.kind == LPC_PC_ABS_EQ;
    if((two && p.n_terms != 2) || (p.kind == LPC_PC_NEQ && p.n_terms > 2)) { set_error("lpc_pc_table_create: propagator %lld has the wrong arity for its kind", (long long)i); return LPC_ERR_INVALID; }
    if(p.kind == LPC_PC_REIF_LIN_LE && (p.bvar < 0 || p.bvar >= nvars)) { set_error("lpc_pc_table_create: propagator %lld has a bad reification variable", (long long)i); return LPC_ERR_INVALID; }
    for(int k = 0; k < p.n_terms; ++k) {
      const lpc_pc_term& t = terms[p.first_term + k];
      if(t.var < 0 || t.var >= nvars) { set_error("lpc_pc_table_create: propagator %lld has a variable out of range", (long long)i); return LPC_ERR_INVALID; }
      if((p.kind == LPC_PC_LIN_LE || p.kind == LPC_PC_REIF_LIN_LE || p.kind == LPC_PC_CLAUSE) && t.coef == 0) { set_error("lpc_pc_table_create: propagator %lld has a zero coefficient", (long long)i); return LPC_ERR_INVALID; }
    }
    hdr[i] = make_int4(p.kind | (p.n_terms << 8), p.first_term, p.rhs, p.bvar);
    has_linear |= p.kind == LPC_PC_LIN_LE || p.kind == LPC_PC_REIF_LIN_LE;
  }
  lpc_pc_table* t = new lpc_pc_table();
  t->has_linear = has_linear;
  LPC_CUDA(cudaMalloc(&t->d_hdr, hdr.size() * sizeof(int4)));
  LPC_CUDA(cudaMalloc(&t->d_terms, std::max<size_t>((size_t)n_terms * sizeof(int2), 16)));
  if(n_props) LPC_CUDA(cudaMemcpy(t->d_hdr, hdr.data(), (size_t)n_props * sizeof(int4), cudaMemcpyHostToDevice));
  if(n_terms) LPC_CUDA(cudaMemcpy(t->d_terms, terms, (size_t)n_terms * sizeof(int2), cudaMemcpyHostToDevice));
  t->dev.hdr = (const int4*)t->d_hdr; t->dev.terms = (const int2*)t->d_terms;
  t->dev.n = n_props; t->dev.n_terms = n_terms; t->dev.nvars = nvars;
  int dev = 0;
  LPC_CUDA(cudaGetDevice(&dev));
  LPC_CUDA(cudaDeviceGetAttribute(&t->sm_count, cudaDevAttrMultiProcessorCount, dev));
  LPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&t->blocks_per_sm[0], k_pc_fixpoint<false>, PC_TPB, 0));
  LPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&t->blocks_per_sm[1], k_pc_fixpoint<true>, PC_TPB, 0));
  *out = t;
  return LPC_OK;
}

int lpc_pc_table_destroy(lpc_pc_table* t) {
  if(!t) return LPC_OK;
  cudaFree(t->d_hdr); cudaFree(t->d_terms);
  if(t->host_store) lpc_store_destroy(t->host_store);
  delete t;
  return LPC_OK;
}

int64_t lpc_pc_table_size(const lpc_pc_table* t) { return t ? (int64_t)t->dev.n : 0; }
int64_t lpc_pc_table_terms(const lpc_pc_table* t) { return t ? (int64_t)t->dev.n_terms : 0; }

static int check_bits(const lpc_pc_table* t, bool bits) {
  if(bits && t->has_linear) {
    set_error("bitset stores support the EQ, NEQ, CLAUSE and ABS_EQ kinds only (NBitset arithmetic is unpinned upstream)");
    return LPC_ERR_UNSUPPORTED;
  }
  return LPC_OK;
}

static int pc_fixpoint_async(const lpc_pc_table* t, lpc_store* s, const lpc_fixpoint_opts* o, bool bits) {
  LPC_REQUIRE(t && s, "null argument");
  LPC_REQUIRE(s->nvars >= t->dev.nvars, "store smaller than the table's variable range");
  LPC_REQUIRE(t->blocks_per_sm[bits] > 0, "kernel does not fit on an SM");
  if(int rc = check_bits(t, bits)) return rc;
  lpc_fixpoint_opts def;
  if(!o) { lpc_fixpoint_default_opts(&def); o = &def; }
  cudaStream_t st = (cudaStream_t)o->stream;
  int grid = t->sm_count * t->blocks_per_sm[bits];
  long long want = std::max<long long>(1, (t->dev.n + PC_TPB - 1) / PC_TPB);
  if(want < grid) grid = (int)want;
  LPC_CUDA(cudaEventRecord(s->ev0, st));
  LPC_CUDA(cudaMemsetAsync(s->d_ctl, 0, sizeof(FixCtl), st));
  PcTableDev td = t->dev;
  int2* store = s->d;
  FixCtl* ctl = s->d_ctl;
  int max_sweeps = o->max_sweeps, stop = o->stop_on_bot;
  void* args[] = {&td, &store, &ctl, &max_sweeps, &stop};
  LPC_CUDA(cudaLaunchCooperativeKernel(bits ? (void*)k_pc_fixpoint<true> : (void*)k_pc_fixpoint<false>, dim3(grid),
                                       dim3(PC_TPB), args, 0, st));
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
  if(int rc = check_bits(t, bits)) return rc;
  if(bits) k_pc_deduce_one<true><<<1, 1>>>(t->dev, s->d, i, &s->d_ctl->scratch[0]);
  else k_pc_deduce_one<false><<<1, 1>>>(t->dev, s->d, i, &s->d_ctl->scratch[0]);
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
  if(int rc = check_bits(t, bits)) return rc;
  unsigned long long* d_cnt = nullptr;
  uint8_t* d_bits = nullptr;
  LPC_CUDA(cudaMalloc((void**)&d_cnt, 8));
  LPC_CUDA(cudaMemset(d_cnt, 0, 8));
  if(bits_out && t->dev.n) LPC_CUDA(cudaMalloc((void**)&d_bits, t->dev.n));
  if(t->dev.n) {
    int blocks = (int)std::min<long long>(ceil_div(t->dev.n, 256), 148 * 8);
    if(bits) k_pc_ask_all<true><<<blocks, 256>>>(t->dev, s->d, d_cnt, d_bits);
    else k_pc_ask_all<false><<<blocks, 256>>>(t->dev, s->d, d_cnt, d_bits);
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
