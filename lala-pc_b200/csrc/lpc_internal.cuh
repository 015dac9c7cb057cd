// lpc_internal.cuh — handle layouts and helpers shared by the translation units of liblpc.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include <atomic>

#include "../../include/lpc.h"
#include "pir_device.cuh"

namespace lpc {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
extern std::atomic<int64_t> g_launches;

#define LPC_CUDA(call)                                                            \
  do {                                                                            \
    cudaError_t e__ = (call);                                                     \
    if(e__ != cudaSuccess) return lpc::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while(0)

#define LPC_REQUIRE(cond, msg)                  \
  do {                                          \
    if(!(cond)) {                               \
      lpc::set_error("%s: %s", __func__, msg);  \
      return LPC_ERR_INVALID;                   \
    }                                           \
  } while(0)

// Device image of the propagator table: opcode bytes + x/y/z index arrays (SoA, 13 B per record), padded with
// D_NOP records to a multiple of 4 so that a thread fetches 4 records with one 32-bit and three 128-bit loads.
struct TableDev {
  const uint8_t* op;
  const int* x;
  const int* y;
  const int* z;
  long long n;       // records
  long long n_pad;   // multiple of 4
  int nvars;
  // var -> records incidence (CSR), for the change-driven worklist
  const int* inc_off;   // [nvars + 1]
  const int* inc_idx;   // [3 n] (duplicates removed per record)
};

// Runs of equal opcode in table order (exact record indices, padding excluded). A table built by PIR::deduce(tell) is
// sorted by (op, y, x, z) (pir.hpp:343-347) and has at most ten runs; n == 0 means "more than LPC_MAX_OPSEG runs", i.e.
// not sorted by opcode, and the kernels fall back to dispatching per record.
#define LPC_MAX_OPSEG 16
struct OpSegs {
  int n;
  int start[LPC_MAX_OPSEG + 1];
  unsigned char op[LPC_MAX_OPSEG];
};

// Control block of one fixpoint run (device memory, one per store handle).
struct FixCtl {
  int flags[4];         // rotating per-iteration flag words: bit0 = changed, bit1 = bot
  int sweeps;
  int dense_sweeps;
  int has_changed;
  int is_bot;
  unsigned long long deductions;
  int hazard;           // a finite bound next to the int32 limits was seen (lpc.h: overflow_hazard)
  int pad_[3];
  int scratch[4];
  unsigned long long bar[4];   // rotating vote-carrying barrier words (grid_barrier.cuh; 3 in use)
  int sm_slots[256];    // blocks arrived per SM (sm_rank_arrive / sm_rank_resolve, grid_barrier.cuh)
};

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

} // namespace lpc

struct lpc_store;
int lpc_dirty_fixpoint_launch(struct lpc_table* t, lpc_store* s, const lpc_fixpoint_opts* o);   // pir_dirty.cu
int lpc_check_device(int handle_device, const char* what);                                       // pir_fixpoint.cu
extern "C" int lpc_table_ensure_csr(struct lpc_table* t);                                                   // lpc_core.cu
struct lpc_table {
  int device = 0;
  std::vector<lpc_bytecode> host;   // the caller's records, caller's order (load_deduce)
  lpc::TableDev dev{};
  void* d_op = nullptr; void* d_x = nullptr; void* d_y = nullptr; void* d_z = nullptr;
  void* d_inc_off = nullptr; void* d_inc_idx = nullptr;
  bool has_div = false;
  long long op_count[10] = {0};
  // incremental build (lpc_table_append / lpc_table_finalize)
  long long cap_pad = 0;            // records the device arrays have room for
  long long dirty_from = 0;         // first record whose device image is stale
  long long sorted_n = 0;           // the first sorted_n records are in (op, y, x, z) order
  long long uploaded_bytes = 0;     // table bytes copied to the device so far (diagnostic: lpc_table_uploaded_bytes)
  bool finalized = false;
  bool csr_valid = false;           // var -> records index (lpc_table_ensure_csr)
  long long generation = 0;         // bumped by every finalize; batches remember the one they were created over
  lpc::OpSegs opsegs{};             // opcode runs for the per-operator loops of the block kernels (pir_batch.cu)
  // launch plan of the dense sweep, computed once on first use (opcode segments in quads, blocks per SM)
  bool plan_ready = false;
  int seg_n = 0;
  int seg_q[17] = {0};
  int blocks_per_sm[2] = {0, 0};    // [track]
  int sm_count = 0;
  size_t smem_optin = 0;            // cudaDevAttrMaxSharedMemoryPerBlockOptin
  lpc_store* host_store = nullptr;  // device staging store of lpc_fixpoint_host
  bool dirty_ready = false;         // occupancy of the change-driven kernel (pir_dirty.cu)
  int dirty_blocks_per_sm[2] = {0, 0};
};

struct lpc_store {
  int device = 0;
  int nvars = 0;
  int2* d = nullptr;
  bool owning = true;
  lpc::FixCtl* d_ctl = nullptr;
  lpc::FixCtl* h_ctl = nullptr;     // pinned mirror
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStream_t last_stream = nullptr;
  bool pending = false;
  // group byte maps of the change-driven kernel (pir_dirty.cu)
  unsigned char* d_dirty = nullptr; long long dirty_cap = 0;
  // the operand cells each lane tile of the PC kernel saw at its last evaluation (pc_fixpoint.cu, LPC_MODE_AUTO)
  int2* d_pc_seen = nullptr; long long pc_seen_cap = 0;
};
