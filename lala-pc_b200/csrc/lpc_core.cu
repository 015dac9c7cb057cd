// lpc_core.cu — library plumbing: errors, device selection, propagator table upload, interval store.
#include "lpc_internal.cuh"
#include "../../include/lpc_pc.h"

#include <cstdarg>
#include <cstring>
#include <algorithm>
#include <climits>

namespace lpc {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
  return e == cudaErrorMemoryAllocation ? LPC_ERR_NOMEM : LPC_ERR_CUDA;
}

// lala-core Sig -> dense device opcode.
static int to_dev_op(int sig) {
  switch(sig) {
    case LPC_ADD: return D_ADD; case LPC_MUL: return D_MUL; case LPC_MIN: return D_MIN; case LPC_MAX: return D_MAX;
    case LPC_TDIV: return D_TDIV; case LPC_FDIV: return D_FDIV; case LPC_CDIV: return D_CDIV; case LPC_EDIV: return D_EDIV;
    case LPC_EQ: return D_EQ; case LPC_LEQ: return D_LEQ;
    default: return -1;
  }
}

// ---- small store kernels -------------------------------------------------------------------------------------------
__global__ void k_store_fill_top(int2* s, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) s[i] = make_int2(LPC_MINF, LPC_INF);
}

// out[0] |= any empty, out[1] |= any non-top
__global__ void k_store_scan(const int2* s, int n, int* out) {
  int bot = 0, nontop = 0;
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int2 v = s[i];
    bot |= v.x > v.y;
    nontop |= (v.x != LPC_MINF) | (v.y != LPC_INF);
  }
  bot = __syncthreads_or(bot);
  nontop = __syncthreads_or(nontop);
  if(threadIdx.x == 0) {
    if(bot) atomicOr(&out[0], 1);
    if(nontop) atomicOr(&out[1], 1);
  }
}

// the same over NBitset<64> cells: empty = 0, top = all ones
__global__ void k_store_scan_bits(const unsigned long long* s, int n, int* out) {
  int bot = 0, nontop = 0;
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned long long v = s[i];
    bot |= v == 0;
    nontop |= v != ~0ull;
  }
  bot = __syncthreads_or(bot);
  nontop = __syncthreads_or(nontop);
  if(threadIdx.x == 0) {
    if(bot) atomicOr(&out[0], 1);
    if(nontop) atomicOr(&out[1], 1);
  }
}
__global__ void k_store_embed_bits(unsigned long long* s, int var, unsigned long long cell, int* changed) {
  const unsigned long long v = s[var], n = v & cell;
  if(n != v) s[var] = n;
  *changed = n != v;
}

// VStore::embed: meet + changed
__global__ void k_store_embed(int2* s, int var, int lb, int ub, int* changed) {
  int2 v = s[var];
  int c = 0;
  if(lb > v.x) { v.x = lb; c = 1; }
  if(ub < v.y) { v.y = ub; c = 1; }
  if(c) s[var] = v;
  *changed = c;
}

// pir.hpp:333-335
__global__ void k_clamp_reified(TableDev t, int2* s) {
  for(long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < t.n; i += (long long)gridDim.x * blockDim.x) {
    int op = t.op[i];
    if(op == D_EQ || op == D_LEQ) {
      int x = t.x[i];
      atomicMax(&s[x].x, 0);
      atomicMin(&s[x].y, 1);
    }
  }
}

} // namespace lpc

using namespace lpc;

// grid-stride 128-bit copy (lpc_measure_l2_copy_gbs)
__global__ void k_copy16(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n) {
  for(size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

extern "C" {

const char* lpc_version(void) { return "lpc-b200 0.1 (sm_100a)"; }
const char* lpc_last_error(void) { return g_err; }
int64_t lpc_launch_count(void) { return g_launches.load(); }

int lpc_measure_l2_copy_gbs(int64_t bytes, int iters, double* gbs) {
  LPC_REQUIRE(gbs != nullptr && bytes >= 4096 && iters >= 1, "bad argument");
  const size_t n16 = (size_t)bytes / 16;
  uint4 *src = nullptr, *dst = nullptr;
  LPC_CUDA(cudaMalloc((void**)&src, n16 * 16));
  LPC_CUDA(cudaMalloc((void**)&dst, n16 * 16));
  LPC_CUDA(cudaMemset(src, 1, n16 * 16));
  int dev = 0, sms = 0;
  LPC_CUDA(cudaGetDevice(&dev));
  LPC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cudaEvent_t e0, e1;
  LPC_CUDA(cudaEventCreate(&e0));
  LPC_CUDA(cudaEventCreate(&e1));
  const int grid = sms * 8;
  for(int i = 0; i < 2; ++i) k_copy16<<<grid, 512>>>(src, dst, n16);
  LPC_CUDA(cudaEventRecord(e0));
  for(int i = 0; i < iters; ++i) k_copy16<<<grid, 512>>>(src, dst, n16);
  LPC_CUDA(cudaEventRecord(e1));
  LPC_CUDA(cudaEventSynchronize(e1));
  float ms = 0;
  LPC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  g_launches += iters + 2;
  *gbs = 2.0 * (double)(n16 * 16) * iters / ((double)ms * 1e-3) / 1e9;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(src); cudaFree(dst);
  return LPC_OK;
}

int lpc_device_count(int* out) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if(e != cudaSuccess) { n = 0; cudaGetLastError(); }
  if(out) *out = n;
  return LPC_OK;
}

int lpc_device_init(int device) {
  int n = 0;
  lpc_device_count(&n);
  if(n == 0) {
    set_error("no CUDA device: this library has no CPU path");
    return LPC_ERR_NO_DEVICE;
  }
  LPC_REQUIRE(device >= 0 && device < n, "device index out of range");
  LPC_CUDA(cudaSetDevice(device));
  LPC_CUDA(cudaFree(0));
  return LPC_OK;
}

// ---- table ---------------------------------------------------------------------------------------------------------
// The host keeps the records in the caller's order (load_deduce); the device keeps the SoA image with spare capacity, so
// that an incremental tell (lpc_table_append + lpc_table_finalize) uploads only the part of the arrays that changed.
static int no_device() {
  int cnt = 0;
  lpc_device_count(&cnt);
  if(cnt == 0) { set_error("no CUDA device: this library has no CPU path"); return LPC_ERR_NO_DEVICE; }
  return LPC_OK;
}

int lpc_table_create_empty(int32_t nvars, lpc_table** out) {
  LPC_REQUIRE(out != nullptr, "null out");
  LPC_REQUIRE(nvars >= 0, "bad nvars");
  int rc = no_device();
  if(rc) return rc;
  lpc_table* t = new lpc_table();
  cudaError_t e = cudaGetDevice(&t->device);
  if(e != cudaSuccess) { delete t; return cuda_fail(e, "cudaGetDevice", __FILE__, __LINE__); }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, t->device);
  if(e != cudaSuccess) { delete t; return cuda_fail(e, "cudaGetDeviceProperties", __FILE__, __LINE__); }
  t->sm_count = prop.multiProcessorCount;
  t->smem_optin = prop.sharedMemPerBlockOptin;
  t->dev.nvars = nvars;
  if((rc = lpc_table_finalize(t, 0))) { lpc_table_destroy(t); return rc; }
  *out = t;
  return LPC_OK;
}

int lpc_table_append(lpc_table* t, const lpc_bytecode* records, int64_t n) {
  LPC_REQUIRE(t != nullptr, "null table");
  LPC_REQUIRE(n >= 0 && (n == 0 || records != nullptr), "bad records");
  // 3 incidences per record must fit the int offsets of the var -> records index
  LPC_REQUIRE((int64_t)t->host.size() + n <= (int64_t)(INT_MAX / 3) - 16, "too many records");
  const int nvars = t->dev.nvars;
  for(int64_t i = 0; i < n; ++i) {
    const lpc_bytecode& b = records[i];
    if(to_dev_op(b.op) < 0) { set_error("lpc_table_append: record %lld has unsupported op %d", (long long)i, b.op); return LPC_ERR_UNSUPPORTED; }
    if(b.x < 0 || b.x >= nvars || b.y < 0 || b.y >= nvars || b.z < 0 || b.z >= nvars) {
      set_error("lpc_table_append: record %lld has a variable out of range", (long long)i); return LPC_ERR_INVALID;
    }
  }
  t->dirty_from = std::min<long long>(t->dirty_from, (long long)t->host.size());
  t->host.insert(t->host.end(), records, records + n);
  t->finalized = false;
  return LPC_OK;
}

int lpc_table_set_nvars(lpc_table* t, int32_t nvars) {
  LPC_REQUIRE(t != nullptr, "null table");
  LPC_REQUIRE(nvars >= t->dev.nvars, "the variable range of a table can only grow");
  if(nvars != t->dev.nvars) {
    t->dev.nvars = nvars;
    t->finalized = false;
  }
  return LPC_OK;
}

int lpc_table_truncate(lpc_table* t, int64_t n) {
  LPC_REQUIRE(t != nullptr && n >= 0 && n <= (int64_t)t->host.size(), "bad record count");
  if(n < (int64_t)t->host.size()) {
    t->host.resize((size_t)n);
    t->dirty_from = std::min<long long>(t->dirty_from, n);
    t->finalized = false;
  }
  return LPC_OK;
}

static bool rec_less(const lpc_bytecode& a, const lpc_bytecode& b) {   // pir.hpp:343-347
  if(a.op != b.op) return a.op < b.op;
  if(a.y != b.y) return a.y < b.y;
  if(a.x != b.x) return a.x < b.x;
  return a.z < b.z;
}

int lpc_table_finalize(lpc_table* t, int32_t sort) {
  LPC_REQUIRE(t != nullptr, "null table");
  int rc = lpc_check_device(t->device, "lpc_table_finalize");
  if(rc) return rc;
  const long long n = (long long)t->host.size();
  if(sort) {
    // stable sort by (op, y, x, z). When the part already on the device is sorted, only the appended tail is sorted and
    // merged in, and the image changes from the place where the smallest new record lands.
    const long long old_n = std::min<long long>(t->sorted_n, n);
    if(old_n < n) {
      auto b0 = t->host.begin();
      t->dirty_from = std::min<long long>(t->dirty_from, old_n);   // the unsorted tail may hold records already uploaded
      std::stable_sort(b0 + old_n, t->host.end(), rec_less);
      if(old_n > 0) {
        const long long ins = std::upper_bound(b0, b0 + old_n, t->host[(size_t)old_n], rec_less) - b0;
        t->dirty_from = std::min<long long>(t->dirty_from, ins);
        std::inplace_merge(b0, b0 + old_n, t->host.end(), rec_less);
      }
    }
    t->sorted_n = n;
  }
  else t->sorted_n = std::min<long long>(t->sorted_n, t->dirty_from);
  long long n_pad = (n + 15) / 16 * 16;   // quads for the 128-bit loads, 16-B granules for the bulk copies
  if(n_pad == 0) n_pad = 16;
  LPC_CUDA(cudaDeviceSynchronize());   // no kernel may still be reading the arrays that are about to change
  long long from = std::min<long long>(t->dirty_from, n) / 16 * 16;
  if(n_pad > t->cap_pad) {   // grow geometrically; everything is uploaded again
    const long long cap = std::max<long long>(n_pad, t->cap_pad * 2);
    cudaFree(t->d_op); cudaFree(t->d_x); cudaFree(t->d_y); cudaFree(t->d_z);
    t->d_op = t->d_x = t->d_y = t->d_z = nullptr;
    t->cap_pad = 0;
    LPC_CUDA(cudaMalloc(&t->d_op, (size_t)cap));
    LPC_CUDA(cudaMalloc(&t->d_x, (size_t)cap * 4));
    LPC_CUDA(cudaMalloc(&t->d_y, (size_t)cap * 4));
    LPC_CUDA(cudaMalloc(&t->d_z, (size_t)cap * 4));
    t->cap_pad = cap;
    from = 0;
  }
  {   // upload records [from, n_pad) of the SoA image (NOP padding behind n)
    const long long m = n_pad - from;
    std::vector<uint8_t> op((size_t)m, (uint8_t)D_NOP);
    std::vector<int> x((size_t)m, 0), y((size_t)m, 0), z((size_t)m, 0);
    for(long long i = from; i < n; ++i) {
      const lpc_bytecode& b = t->host[(size_t)i];
      op[i - from] = (uint8_t)to_dev_op(b.op); x[i - from] = b.x; y[i - from] = b.y; z[i - from] = b.z;
    }
    if(m > 0) {
      LPC_CUDA(cudaMemcpy((uint8_t*)t->d_op + from, op.data(), (size_t)m, cudaMemcpyHostToDevice));
      LPC_CUDA(cudaMemcpy((int*)t->d_x + from, x.data(), (size_t)m * 4, cudaMemcpyHostToDevice));
      LPC_CUDA(cudaMemcpy((int*)t->d_y + from, y.data(), (size_t)m * 4, cudaMemcpyHostToDevice));
      LPC_CUDA(cudaMemcpy((int*)t->d_z + from, z.data(), (size_t)m * 4, cudaMemcpyHostToDevice));
    }
    t->uploaded_bytes += (long long)m * 13;
  }
  // operator statistics and opcode runs (OpSegs)
  t->has_div = false;
  for(int d = 0; d < 10; ++d) t->op_count[d] = 0;
  OpSegs& sg = t->opsegs;
  sg.n = 0;
  bool ok = true;
  int prev = -1;
  for(long long i = 0; i < n; ++i) {
    const int d = to_dev_op(t->host[(size_t)i].op);
    t->op_count[d]++;
    if(d >= D_TDIV && d <= D_EDIV) t->has_div = true;
    if(ok && d != prev) {
      if(sg.n == LPC_MAX_OPSEG) ok = false;
      else { sg.op[sg.n] = (unsigned char)d; sg.start[sg.n++] = (int)i; }
    }
    prev = d;
  }
  if(!ok) sg.n = 0;
  sg.start[sg.n] = (int)n;
  t->dev.op = (const uint8_t*)t->d_op; t->dev.x = (const int*)t->d_x; t->dev.y = (const int*)t->d_y; t->dev.z = (const int*)t->d_z;
  t->dev.n = n; t->dev.n_pad = n_pad;
  // whatever was derived from the old contents is rebuilt on first use: the var -> records index, the launch plans, the
  // staging store of lpc_fixpoint_host; batches created over the old contents refuse to run (generation)
  t->csr_valid = false;
  t->dev.inc_off = nullptr; t->dev.inc_idx = nullptr;
  t->plan_ready = false;
  if(t->host_store && lpc_store_nvars(t->host_store) != t->dev.nvars) { lpc_store_destroy(t->host_store); t->host_store = nullptr; }
  t->dirty_from = n;
  t->finalized = true;
  t->generation++;
  return LPC_OK;
}

// var -> records incidence (CSR) for the change-driven kernels, built when one of them first needs it.
int lpc_table_ensure_csr(lpc_table* t) {
  if(t->csr_valid) return LPC_OK;
  LPC_REQUIRE(t->finalized, "lpc_table_finalize has not been called since the last change of the table");
  const long long n = t->dev.n;
  const int nvars = t->dev.nvars;
  std::vector<int> off((size_t)nvars + 1, 0);
  auto each_var = [&](long long i, auto f) {
    const lpc_bytecode& b = t->host[(size_t)i];
    f(b.x);
    if(b.y != b.x) f(b.y);
    if(b.z != b.x && b.z != b.y) f(b.z);
  };
  for(long long i = 0; i < n; ++i) each_var(i, [&](int v) { off[v + 1]++; });
  for(int v = 0; v < nvars; ++v) off[v + 1] += off[v];
  std::vector<int> idx((size_t)std::max(1, off[nvars]));
  {
    std::vector<int> cur(off.begin(), off.end() - 1);
    for(long long i = 0; i < n; ++i) each_var(i, [&](int v) { idx[cur[v]++] = (int)i; });
  }
  LPC_CUDA(cudaDeviceSynchronize());
  cudaFree(t->d_inc_off); cudaFree(t->d_inc_idx);
  t->d_inc_off = t->d_inc_idx = nullptr;
  LPC_CUDA(cudaMalloc(&t->d_inc_off, off.size() * 4));
  LPC_CUDA(cudaMalloc(&t->d_inc_idx, idx.size() * 4));
  LPC_CUDA(cudaMemcpy(t->d_inc_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
  LPC_CUDA(cudaMemcpy(t->d_inc_idx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice));
  t->dev.inc_off = (const int*)t->d_inc_off; t->dev.inc_idx = (const int*)t->d_inc_idx;
  t->csr_valid = true;
  return LPC_OK;
}

int lpc_table_create(const lpc_bytecode* records, int64_t n, int32_t nvars, lpc_table** out) {
  LPC_REQUIRE(out != nullptr, "null out");
  LPC_REQUIRE(n >= 0 && (n == 0 || records != nullptr), "bad records");
  lpc_table* t = nullptr;
  int rc = lpc_table_create_empty(nvars, &t);
  if(rc) return rc;
  if((rc = lpc_table_append(t, records, n)) || (rc = lpc_table_finalize(t, 0))) { lpc_table_destroy(t); return rc; }
  *out = t;
  return LPC_OK;
}

int64_t lpc_table_uploaded_bytes(const lpc_table* t) { return t ? (int64_t)t->uploaded_bytes : 0; }

int lpc_table_destroy(lpc_table* t) {
  if(!t) return LPC_OK;
  cudaFree(t->d_op); cudaFree(t->d_x); cudaFree(t->d_y); cudaFree(t->d_z);
  cudaFree(t->d_inc_off); cudaFree(t->d_inc_idx);
  if(t->host_store) lpc_store_destroy(t->host_store);
  delete t;
  return LPC_OK;
}

int64_t lpc_table_size(const lpc_table* t) { return t ? (int64_t)t->dev.n : 0; }
int32_t lpc_table_nvars(const lpc_table* t) { return t ? t->dev.nvars : 0; }

int lpc_table_load(const lpc_table* t, int64_t i, lpc_bytecode* out) {
  LPC_REQUIRE(t && out, "null argument");
  LPC_REQUIRE(i >= 0 && i < (int64_t)t->host.size(), "record index out of range");
  *out = t->host[i];
  return LPC_OK;
}

int lpc_table_clamp_reified(const lpc_table* t, lpc_store* s) {
  LPC_REQUIRE(t && s, "null argument");
  LPC_REQUIRE(s->nvars >= t->dev.nvars, "store smaller than the table's variable range");
  if(t->dev.n == 0) return LPC_OK;
  int blocks = std::min<long long>(ceil_div(t->dev.n, 256), 148 * 8);
  k_clamp_reified<<<blocks, 256>>>(t->dev, s->d);
  g_launches++;
  LPC_CUDA(cudaGetLastError());
  LPC_CUDA(cudaDeviceSynchronize());
  return LPC_OK;
}

// ---- store ---------------------------------------------------------------------------------------------------------
static int store_init_common(lpc_store* s) {
  LPC_CUDA(cudaMalloc((void**)&s->d_ctl, sizeof(FixCtl)));
  LPC_CUDA(cudaMemset(s->d_ctl, 0, sizeof(FixCtl)));
  LPC_CUDA(cudaHostAlloc((void**)&s->h_ctl, sizeof(FixCtl), cudaHostAllocDefault));
  memset(s->h_ctl, 0, sizeof(FixCtl));
  LPC_CUDA(cudaEventCreate(&s->ev0));
  LPC_CUDA(cudaEventCreate(&s->ev1));
  return LPC_OK;
}

int lpc_store_create(int32_t nvars, lpc_store** out) {
  LPC_REQUIRE(out != nullptr && nvars >= 0, "bad argument");
  int rc = no_device();
  if(rc) return rc;
  lpc_store* s = new lpc_store();
  s->nvars = nvars;
  s->owning = true;
  // every failure below destroys the half-built handle (and what it already allocated) before returning
  auto body = [&]() -> int {
    LPC_CUDA(cudaGetDevice(&s->device));
    LPC_CUDA(cudaMalloc((void**)&s->d, std::max<size_t>((size_t)nvars * 8, 16)));
    if(nvars) {
      k_store_fill_top<<<ceil_div(nvars, 256), 256>>>(s->d, nvars);
      g_launches++;
      LPC_CUDA(cudaGetLastError());
    }
    int rc2 = store_init_common(s);
    if(rc2) return rc2;
    LPC_CUDA(cudaDeviceSynchronize());
    return LPC_OK;
  };
  if((rc = body())) { lpc_store_destroy(s); return rc; }
  *out = s;
  return LPC_OK;
}

int lpc_store_wrap_device(void* device_ptr, int32_t nvars, lpc_store** out) {
  LPC_REQUIRE(out != nullptr && nvars >= 0 && device_ptr != nullptr, "bad argument");
  LPC_REQUIRE(((uintptr_t)device_ptr & 7) == 0, "device pointer must be 8-byte aligned");
  lpc_store* s = new lpc_store();
  s->nvars = nvars;
  s->owning = false;
  s->d = (int2*)device_ptr;
  cudaError_t e = cudaGetDevice(&s->device);
  if(e != cudaSuccess) { lpc_store_destroy(s); return cuda_fail(e, "cudaGetDevice", __FILE__, __LINE__); }
  int rc = store_init_common(s);
  if(rc) { lpc_store_destroy(s); return rc; }
  *out = s;
  return LPC_OK;
}

int lpc_store_destroy(lpc_store* s) {
  if(!s) return LPC_OK;
  if(s->owning) cudaFree(s->d);
  cudaFree(s->d_dirty);
  cudaFree(s->d_pc_seen);
  cudaFree(s->d_ctl);
  if(s->h_ctl) cudaFreeHost(s->h_ctl);
  if(s->ev0) cudaEventDestroy(s->ev0);
  if(s->ev1) cudaEventDestroy(s->ev1);
  delete s;
  return LPC_OK;
}

int32_t lpc_store_nvars(const lpc_store* s) { return s ? s->nvars : 0; }
void* lpc_store_device_ptr(lpc_store* s) { return s ? (void*)s->d : nullptr; }

int lpc_store_write(lpc_store* s, int32_t first, int32_t n, const int32_t* lbub) {
  LPC_REQUIRE(s && (n == 0 || lbub), "null argument");
  LPC_REQUIRE(first >= 0 && n >= 0 && (long long)first + n <= s->nvars, "range out of bounds");
  if(n) LPC_CUDA(cudaMemcpy(s->d + first, lbub, (size_t)n * 8, cudaMemcpyHostToDevice));
  return LPC_OK;
}

int lpc_store_read(const lpc_store* s, int32_t first, int32_t n, int32_t* lbub) {
  LPC_REQUIRE(s && (n == 0 || lbub), "null argument");
  LPC_REQUIRE(first >= 0 && n >= 0 && (long long)first + n <= s->nvars, "range out of bounds");
  if(n) LPC_CUDA(cudaMemcpy(lbub, s->d + first, (size_t)n * 8, cudaMemcpyDeviceToHost));
  return LPC_OK;
}

int lpc_store_embed(lpc_store* s, int32_t var, int32_t lb, int32_t ub, int* changed) {
  LPC_REQUIRE(s != nullptr, "null store");
  LPC_REQUIRE(var >= 0 && var < s->nvars, "variable out of range");
  k_store_embed<<<1, 1>>>(s->d, var, lb, ub, &s->d_ctl->scratch[0]);
  g_launches++;
  LPC_CUDA(cudaGetLastError());
  int c = 0;
  LPC_CUDA(cudaMemcpy(&c, &s->d_ctl->scratch[0], sizeof(int), cudaMemcpyDeviceToHost));
  if(changed) *changed = c;
  return LPC_OK;
}

int lpc_store_copy(lpc_store* dst, const lpc_store* src) {
  LPC_REQUIRE(dst && src, "null argument");
  LPC_REQUIRE(dst->nvars == src->nvars, "stores differ in size");
  if(src->nvars) LPC_CUDA(cudaMemcpy(dst->d, src->d, (size_t)src->nvars * 8, cudaMemcpyDeviceToDevice));
  return LPC_OK;
}

static int store_scan(const lpc_store* s, int out[2], bool bits = false) {
  out[0] = out[1] = 0;
  if(s->nvars == 0) return LPC_OK;
  LPC_CUDA(cudaMemset(&s->d_ctl->scratch[0], 0, 2 * sizeof(int)));
  int blocks = std::min(ceil_div(s->nvars, 256), 148 * 8);
  if(bits) k_store_scan_bits<<<blocks, 256>>>(reinterpret_cast<const unsigned long long*>(s->d), s->nvars, &s->d_ctl->scratch[0]);
  else k_store_scan<<<blocks, 256>>>(s->d, s->nvars, &s->d_ctl->scratch[0]);
  g_launches++;
  LPC_CUDA(cudaGetLastError());
  LPC_CUDA(cudaMemcpy(out, &s->d_ctl->scratch[0], 2 * sizeof(int), cudaMemcpyDeviceToHost));
  return LPC_OK;
}

int lpc_store_is_bot(const lpc_store* s, int* out) {
  LPC_REQUIRE(s && out, "null argument");
  int r[2];
  int rc = store_scan(s, r);
  if(rc) return rc;
  *out = r[0];
  return LPC_OK;
}

int lpc_store_is_top(const lpc_store* s, int* out) {
  LPC_REQUIRE(s && out, "null argument");
  int r[2];
  int rc = store_scan(s, r);
  if(rc) return rc;
  *out = !r[1];
  return LPC_OK;
}

/* ---- NBitset<64> cells (include/lpc_pc.h) ---- */
int lpc_store_embed_bits(lpc_store* s, int32_t var, uint64_t cell, int* changed) {
  LPC_REQUIRE(s != nullptr, "null store");
  LPC_REQUIRE(var >= 0 && var < s->nvars, "variable out of range");
  k_store_embed_bits<<<1, 1>>>(reinterpret_cast<unsigned long long*>(s->d), var, cell, &s->d_ctl->scratch[0]);
  g_launches++;
  LPC_CUDA(cudaGetLastError());
  int c = 0;
  LPC_CUDA(cudaMemcpy(&c, &s->d_ctl->scratch[0], sizeof(int), cudaMemcpyDeviceToHost));
  if(changed) *changed = c;
  return LPC_OK;
}

int lpc_store_is_bot_bits(const lpc_store* s, int* out) {
  LPC_REQUIRE(s && out, "null argument");
  int r[2];
  int rc = store_scan(s, r, true);
  if(rc) return rc;
  *out = r[0];
  return LPC_OK;
}

int lpc_store_is_top_bits(const lpc_store* s, int* out) {
  LPC_REQUIRE(s && out, "null argument");
  int r[2];
  int rc = store_scan(s, r, true);
  if(rc) return rc;
  *out = !r[1];
  return LPC_OK;
}

} // extern "C"
