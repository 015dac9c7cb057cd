// batch_internal.cuh — handle layout of the batched mode, shared by pir_batch.cu and pir_eps.cu.
#pragma once
#include "lpc_internal.cuh"

#define LPC_MAX_RANKS 64
namespace lpc {

struct BatchCtl {
  long long red[4];            // n_solution, n_bot, n_unknown, best_bound (min)
  long long sweeps_total;
  long long deductions;
  int max_sweeps_seen;
  int next_store;              // dynamic scheduler
  int n_surv;                  // EPS: non-failed stores written to the survivor buffer so far (pir_eps.cu)
  int done_blocks;             // blocks that have finished (the last one writes the all-reduce payload)
  int hazard;                  // a store held a finite bound next to the int32 limits (lpc.h: overflow_hazard)
  int rank, world;             // position of this GPU in the multi-GPU job (lpc_batch_set_rank / lpc_eps_set_rank)
  // The payload of the ONE all-reduce (SUM) of the multi-GPU driver: [0..2] = the three counters, [3 + rank] = this
  // rank's best bound (every other slot 0), so that SUM delivers every rank's bound and MIN is taken on the host.
  long long payload[3 + LPC_MAX_RANKS];
};

} // namespace lpc

#define LPC_BATCH_CHUNKS 8
struct lpc_batch {
  const lpc_table* table = nullptr;
  int n_stores = 0, nvars = 0;
  int2* d = nullptr;
  uint8_t* d_flags = nullptr;
  int* d_sweeps = nullptr;
  int* d_obj = nullptr;
  lpc::BatchCtl* d_ctl = nullptr;
  lpc::BatchCtl* h_ctl = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStream_t last_stream = nullptr;
  bool pending = false;
  int sbytes = 0;
  bool plan_ready[2] = {false, false}; // [dense, change-driven]
  bool table_smem[2] = {false, false};
  size_t smem[2] = {0, 0};
  int threads[2] = {0, 0}, grid[2] = {0, 0};
  int* d_seeds = nullptr;                // lpc_batch_set_seeds: variables on which the stores differ from a fixpoint
  int n_seeds = -1;
  lpc::BatchCtl* h_init = nullptr;   // pinned initial control block
  // pipelined host path (lpc_batch_fixpoint_host): copy-in, compute and copy-out streams, per-chunk control blocks
  cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
  cudaEvent_t e_in[LPC_BATCH_CHUNKS] = {nullptr}, e_k[LPC_BATCH_CHUNKS] = {nullptr};
  lpc::BatchCtl* d_ctl_chunk = nullptr;   // [LPC_BATCH_CHUNKS]
  lpc::BatchCtl* h_ctl_chunk = nullptr;   // pinned, [2 * LPC_BATCH_CHUNKS]: results, then initial values
  int n_chunks = 0;                  // chunks of the call in flight (0 = one launch, h_ctl holds the result)
  // grouped kernel over the packed table (pir_eps.cu): scratch for the packed records, and the root store every image of
  // the batch is a tightening of, when known (lpc_batch_init_split establishes it, lpc_batch_write withdraws it)
  void* d_ptab = nullptr; void* d_phdr = nullptr;
  int2* d_root = nullptr; bool root_valid = false;
  // the decision variables of the last lpc_batch_init_split, while every image still IS the root outside them (until the
  // first fixpoint / write): the first sweep of the grouped kernel may then run on their records only (pir_eps.cu)
  int* d_split_vars = nullptr; int n_split_vars = 0; bool split_fresh = false;
  size_t grp_ptab_bytes = 0; int grp_cap1 = 0;
  int grp_g = -1; size_t grp_smem = 0; int grp_grid = 0;   // plan of the grouped kernel (-1 = not decided, 0 = unusable)
  int rank = 0, world = 1;           // lpc_batch_set_rank
  long long table_gen = 0;           // lpc_table::generation at creation: a table that changed since needs a new batch
};

// pir_eps.cu: the grouped kernel over resident store images (called by batch_launch_range of pir_batch.cu).
// *used = 0 when the batch does not qualify (table too large for shared memory, more than 8191 variables, ...).
int lpc_group_launch_resident(lpc_batch* b, const lpc_fixpoint_opts* o, int32_t objective_var, int first, int count,
                              lpc::BatchCtl* d_ctl, lpc::BatchCtl* h_init, cudaStream_t st, int* used);
