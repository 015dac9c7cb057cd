/* lpc.h — C-ABI of the B200-native propagation engine (drop-in for the fixpoint hot path of lala-pc).
 *
 * Everything here is `extern "C"`, plain pointers and sizes. No torch / C++ types cross this boundary.
 * Each entry point cites the reference interface it replaces (paths relative to the lala-pc repository).
 *
 * Conventions
 *   - every function returns an `int` status (LPC_OK == 0); `lpc_last_error()` gives a message (thread-local).
 *   - a *store* is an array of `nvars` intervals, laid out as interleaved int32 pairs {lb, ub} (8 B per
 *     variable), i.e. the memory image of `VStore<Interval<ZLB>>::data` (used at pir.hpp:724-726, 813-815).
 *     top = [INT32_MIN, INT32_MAX]; an interval with lb > ub is empty (bot).
 *   - a *table* is the immutable propagator table `battery::vector<bytecode_type>` (pir.hpp:104, 115-118).
 *   - arithmetic is 32-bit two's-complement with wrap-around (the reference has UB on overflow,
 *     pir.hpp:759-772; generators keep magnitudes small so that the question never arises). The rules are monotone
 *     over the integers, not over wrapped int32: once a FINITE bound sits next to INT32_MIN / INT32_MAX (it gets there
 *     from an exactly infinite one, e.g. `x = 0 => y >= z.lb + 1`, pir.hpp:750-753) a later `+` or `*` may wrap and the
 *     result then depends on the evaluation order - on the reference's side it is undefined behaviour. The engine
 *     reports that situation instead of hiding it: `overflow_hazard` in the result records is set when a finite bound
 *     within 2^24 of the int32 limits was present in the initial store or written during the fixpoint.
 *   - one in-flight call per handle; distinct handles are independent. Handles live on the CUDA device that
 *     was current when they were created.
 */
#ifndef LPC_H
#define LPC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LPC_OK 0
#define LPC_ERR_INVALID 1   /* bad argument */
#define LPC_ERR_CUDA 2      /* CUDA runtime error, see lpc_last_error() */
#define LPC_ERR_NOMEM 3
#define LPC_ERR_UNSUPPORTED 4
#define LPC_ERR_NO_DEVICE 5 /* no CUDA device: the product has no CPU fallback */

/* Operator codes of `bytecode_type::op`. Numeric values follow lala-core's `Sig` enumeration (v1.2.8,
 * un-vendored: recalled, to be static_assert-ed on the reference side). Only the ten
 * operators PIR accepts (pir.hpp:270-273) are valid. */
enum lpc_sig {
  LPC_ADD = 2, LPC_MUL = 4, LPC_MIN = 6, LPC_MAX = 7,
  LPC_TDIV = 25, LPC_FDIV = 27, LPC_CDIV = 29, LPC_EDIV = 31,
  LPC_EQ = 46, LPC_LEQ = 48
};

/* The constraint `x = y op z`. Mirrors `struct bytecode_type {Sig op; AVar x, y, z;}` (pir.hpp:37-47; 16 B,
 * size asserts pir.hpp:112-113). x, y, z are plain variable indices into the store (`AVar::vid()`). */
typedef struct lpc_bytecode {
  int32_t op;
  int32_t x, y, z;
} lpc_bytecode;

typedef struct lpc_table lpc_table;
typedef struct lpc_store lpc_store;
typedef struct lpc_batch lpc_batch;

/* ---- library ------------------------------------------------------------------------------------------ */
const char* lpc_version(void);
const char* lpc_last_error(void);
/* Select the CUDA device for subsequent handle creation on this thread. Fails with LPC_ERR_NO_DEVICE when no
 * GPU is present (there is deliberately no CPU path in the product). */
int lpc_device_init(int device);
int lpc_device_count(int* out);
/* Number of kernels of this library launched by the calling process so far (bench.py's `gpu_launches`). */
int64_t lpc_launch_count(void);
/* Measurement aid (bench.py's roofline): copy bandwidth of a buffer pair that stays resident in L2 (`bytes` each, e.g.
 * 32 MB), read + written bytes per second in GB/s, from `iters` timed repetitions of a grid-stride 128-bit copy kernel
 * after two warm-up passes. The driver provides an HBM figure only; working sets like config 2's store and table live in
 * L2, so the on-chip ceiling is measured by the build itself. */
int lpc_measure_l2_copy_gbs(int64_t bytes, int iters, double* gbs);

/* ---- propagator table (PIR::deduce(tell) build step, pir.hpp:326-352) ------------------------------------ */
/* Upload `n` records (host AoS, the order the caller wants `deduce(i)` / `ask(i)` to index — the façade
 * sorts by (op, y, x, z) like pir.hpp:343-347 before calling). The device keeps an opcode-byte + x/y/z SoA
 * (13 B per record) read with 128-bit loads, plus a var->records CSR for the change-driven worklist.
 * Validates op codes and 0 <= x,y,z < nvars. The EQ/LEQ result clamp to [0,1] (pir.hpp:333-335) is a store
 * operation and is applied by lpc_table_clamp_reified(). */
int lpc_table_create(const lpc_bytecode* records, int64_t n, int32_t nvars, lpc_table** out);
int lpc_table_destroy(lpc_table* t);
/* Incremental build, the way PIR::deduce(tell) grows its table (pir.hpp:326-352; the reference's tests tell between
 * fixpoints, tests/pir_test.cpp:485, 577-581): an empty table over `nvars` variables; lpc_table_append adds records on the
 * host side; lpc_table_finalize brings the device image up to date - with sort != 0 after the stable sort by
 * (op, y, x, z) of pir.hpp:343-347 (only the appended tail is sorted and merged in), with sort == 0 in append order. The
 * device arrays keep spare capacity (doubling), and only the part of the image behind the first record that moved is
 * uploaded; the var -> records index of the change-driven kernels is rebuilt on first use. lpc_table_set_nvars widens the
 * variable range (a tell that declares variables), lpc_table_truncate keeps the first n records (PIR::restore pops from
 * the back of the table, pir.hpp:863-870). No call using the table may be in flight; batches and EPS handles created over
 * the table before a finalize must be re-created. */
int lpc_table_create_empty(int32_t nvars, lpc_table** out);
int lpc_table_append(lpc_table* t, const lpc_bytecode* records, int64_t n);
int lpc_table_set_nvars(lpc_table* t, int32_t nvars);
int lpc_table_truncate(lpc_table* t, int64_t n);
int lpc_table_finalize(lpc_table* t, int32_t sort);
/* Table bytes copied to the device since the table was created (diagnostic: what incremental tells cost). */
int64_t lpc_table_uploaded_bytes(const lpc_table* t);
/* PIR::num_deductions (pir.hpp:382-384). */
int64_t lpc_table_size(const lpc_table* t);
int32_t lpc_table_nvars(const lpc_table* t);
/* PIR::load_deduce (pir.hpp:358-366): record i, from the host mirror. */
int lpc_table_load(const lpc_table* t, int64_t i, lpc_bytecode* out);
/* pir.hpp:333-335: for every EQ/LEQ record, meet store[x] with [0,1]. */
int lpc_table_clamp_reified(const lpc_table* t, lpc_store* s);

/* ---- interval store (VStore<Interval<ZLB>>) -------------------------------------------------------------- */
int lpc_store_create(int32_t nvars, lpc_store** out);      /* every variable at top */
/* Non-owning view over caller-provided device memory of nvars*8 bytes (e.g. a torch tensor's data_ptr). */
int lpc_store_wrap_device(void* device_ptr, int32_t nvars, lpc_store** out);
int lpc_store_destroy(lpc_store* s);
int32_t lpc_store_nvars(const lpc_store* s);               /* PIR::vars, pir.hpp:853-855 */
void* lpc_store_device_ptr(lpc_store* s);
/* Plain assignment of variables [first, first+n) from host pairs {lb,ub}. */
int lpc_store_write(lpc_store* s, int32_t first, int32_t n, const int32_t* lbub);
/* PIR::operator[] / project (pir.hpp:840-851): copy variables [first, first+n) to host. */
int lpc_store_read(const lpc_store* s, int32_t first, int32_t n, int32_t* lbub);
/* PIR::embed (pir.hpp:354-356): store[var] := store[var] meet [lb,ub]; *changed as VStore::embed. */
int lpc_store_embed(lpc_store* s, int32_t var, int32_t lb, int32_t ub, int* changed);
/* snapshot/restore of the store part (pir.hpp:857-870): device-to-device copy. */
int lpc_store_copy(lpc_store* dst, const lpc_store* src);
/* PIR::is_bot (pir.hpp:831-833): 1 iff some variable is empty. PIR::is_top's store half (pir.hpp:836-838). */
int lpc_store_is_bot(const lpc_store* s, int* out);
int lpc_store_is_top(const lpc_store* s, int* out);

/* ---- the hot path: one fixpoint (GaussSeidelIteration::fixpoint(n, deduce) call sites:
 *      tests/pir_test.cpp:60-62, 82-86; bound_consistency_test.hpp:36-39) ---------------------------------- */
#define LPC_MODE_AUTO 0      /* change-driven: dense sweeps that skip 64-record groups none of whose variables changed, once
                                a sweep changes at most 1/d of the groups (d = opts.reserved, default 8) */
#define LPC_MODE_SWEEP 1     /* dense: every sweep evaluates every propagator */
#define LPC_MODE_WORKLIST 2  /* change-driven from the second sweep on (the same kernel as AUTO, handing over at once) */

typedef struct lpc_fixpoint_opts {
  int32_t mode;          /* LPC_MODE_* */
  int32_t max_sweeps;    /* 0 = unlimited */
  int32_t stop_on_bot;   /* 1 (default contract): stop at the first sweep that observes an empty variable */
  int32_t reserved;      /* LPC_MODE_AUTO: hand-over divisor d (0 = default) */
  uint64_t stream;       /* cudaStream_t to enqueue on (0 = the default stream) */
} lpc_fixpoint_opts;

typedef struct lpc_fixpoint_result {
  int32_t has_changed;   /* the `has_changed` out-parameter of fixpoint(n, f, has_changed) */
  int32_t is_bot;        /* store is at bot (contents then unspecified, like the reference's failed stores) */
  int32_t sweeps;        /* iterations of the outer loop (informational: schedule dependent) */
  int32_t dense_sweeps;  /* how many of them were dense */
  int64_t deductions;    /* deduce() evaluations executed (informational: schedule dependent) */
  float device_ms;       /* device time of the fixpoint kernel(s), CUDA events on `stream` */
  int32_t overflow_hazard; /* 1: a finite bound next to the int32 limits was seen (see "arithmetic" above) */
} lpc_fixpoint_result;

void lpc_fixpoint_default_opts(lpc_fixpoint_opts* o);
/* Run deduce(i) for all i until no bound changes. Store resident on the device. Synchronous. */
int lpc_fixpoint(const lpc_table* t, lpc_store* s, const lpc_fixpoint_opts* o, lpc_fixpoint_result* r);
/* Enqueue only (no host synchronisation); collect later with lpc_fixpoint_collect on the same store. */
int lpc_fixpoint_async(const lpc_table* t, lpc_store* s, const lpc_fixpoint_opts* o);
int lpc_fixpoint_collect(lpc_store* s, lpc_fixpoint_result* r);
/* Same with HOST buffers: copies `lbub` (nvars pairs) to the device, runs the fixpoint, copies it back. */
int lpc_fixpoint_host(const lpc_table* t, int32_t* lbub, const lpc_fixpoint_opts* o, lpc_fixpoint_result* r);

/* PIR::deduce(int i) (pir.hpp:387-390, 721-817): a single propagator step, for façade parity. */
int lpc_deduce_one(const lpc_table* t, lpc_store* s, int64_t i, int* changed);
/* PIR::ask(int i) (pir.hpp:368-370, 417-438). */
int lpc_ask_one(const lpc_table* t, const lpc_store* s, int64_t i, int* entailed);
/* The ask loop of PIR::is_extractable (pir.hpp:873-884): number of entailed records; all = (n == size). */
int lpc_ask_all(const lpc_table* t, const lpc_store* s, int64_t* n_entailed);
/* Per-record entailment bits (one byte per record) to host. */
int lpc_ask_bits(const lpc_table* t, const lpc_store* s, uint8_t* out);

/* ---- batched mode: one store per subproblem, one thread block per store ---------------------------------- */
/* The table is shared by all stores (the `deps.is_shared_copy()` design of pir.hpp:182-195). A store image travels as one
 * bulk copy, whose size must be a multiple of 16 bytes: the table must count an EVEN number of variables
 * (lpc_table_set_nvars adds an unused one), else LPC_ERR_UNSUPPORTED. The same holds for lpc_eps_create. */
int lpc_batch_create(const lpc_table* t, int32_t n_stores, lpc_batch** out);
int lpc_batch_destroy(lpc_batch* b);
void* lpc_batch_device_ptr(lpc_batch* b);                 /* [n_stores][nvars] pairs */
int lpc_batch_write(lpc_batch* b, int32_t first_store, int32_t n, const int32_t* lbub);
int lpc_batch_read(const lpc_batch* b, int32_t first_store, int32_t n, int32_t* lbub);
/* EPS decomposition on the device: store k := base, then for decision j (bit j of first_id + k):
 * var d_j keeps its lower half [lb, mid] (bit 0) or upper half [mid+1, ub] (bit 1), mid = lb + (ub-lb)/2
 * taken on the base domain. */
int lpc_batch_init_split(lpc_batch* b, const int32_t* base_lbub, const int32_t* decision_vars, int32_t n_decisions,
                         int64_t first_id);
/* Same with an explicit subproblem id per store (n_stores ids): lets a multi-GPU driver hand every rank a uniform
 * sample of the id space instead of a sub-cube of it. */
int lpc_batch_init_split_ids(lpc_batch* b, const int32_t* base_lbub, const int32_t* decision_vars, int32_t n_decisions,
                             const int64_t* ids);

/* A promise that makes the change-driven modes incremental: every store of the batch is a fixpoint of the table except
 * possibly on these `n` variables (e.g. the decision variables of an EPS split of a root fixpoint). The first sweep of
 * lpc_batch_fixpoint (LPC_MODE_WORKLIST) and the root node of a change-driven lpc_batch_search then only evaluate the
 * propagators incident to them. n = -1 withdraws the promise (default): the first sweep evaluates every propagator. A false promise gives an
 * under-propagated store; LPC_MODE_SWEEP ignores it. */
int lpc_batch_set_seeds(lpc_batch* b, const int32_t* vars, int32_t n);

typedef struct lpc_batch_result {
  int64_t n_bot;         /* stores that failed */
  int64_t n_solution;    /* non-failed stores on which every propagator is entailed (is_extractable) */
  int64_t n_unknown;     /* the rest */
  int32_t best_bound;    /* min over non-failed stores of lb(objective_var); INT32_MAX if none */
  int32_t max_sweeps_seen;
  int64_t sweeps_total;  /* sum over stores */
  int64_t deductions;    /* deduce() evaluations executed over the whole batch */
  float device_ms;
  int32_t overflow_hazard; /* 1: some store held a finite bound next to the int32 limits before or after its fixpoint */
} lpc_batch_result;

/* Fixpoint of every store of the batch. objective_var < 0: no objective. opts.mode: LPC_MODE_SWEEP and LPC_MODE_AUTO =
 * dense sweeps (every propagator every sweep), LPC_MODE_WORKLIST = change-driven: a 32-record group is re-evaluated only
 * when a variable of one of its records changed (pays off when changes stay local). Same fixpoints either way. */
int lpc_batch_fixpoint(lpc_batch* b, const lpc_fixpoint_opts* o, int32_t objective_var, lpc_batch_result* r);
int lpc_batch_fixpoint_async(lpc_batch* b, const lpc_fixpoint_opts* o, int32_t objective_var);
int lpc_batch_collect(lpc_batch* b, lpc_batch_result* r);
/* Same with HOST buffers ([n_stores][nvars] pairs in, fixpoints out). A large batch (>= 64 MB) goes through in 8 chunks
 * on three streams, so that copy-in, fixpoints and copy-out overlap on the full-duplex link; give it pinned memory for
 * that to happen. Synchronous: the buffer holds the results on return. */
int lpc_batch_fixpoint_host(lpc_batch* b, int32_t* lbub, const lpc_fixpoint_opts* o, int32_t objective_var,
                            lpc_batch_result* r);
/* Per-store flags to host: bit0 = bot, bit1 = all propagators entailed. */
int lpc_batch_flags(const lpc_batch* b, uint8_t* out);
/* Device address of the 4 x int64 reduction record {n_solution, n_bot, n_unknown, best_bound} of the last
 * lpc_batch_fixpoint: the payload of the one NCCL all-reduce of the multi-GPU driver. */
void* lpc_batch_reduction_device_ptr(lpc_batch* b);

/* Multi-GPU: tell the batch its position in the job. The kernel that finishes a lpc_batch_fixpoint then also fills the
 * all-reduce payload (below) - [0..2] the three counters, [3 + rank] this rank's best bound, every other slot 0 - so that
 * ONE all-reduce (SUM) of 3 + world int64 delivers the counters and every rank's bound (MIN is taken on the host). */
int lpc_batch_set_rank(lpc_batch* b, int32_t rank, int32_t world);
void* lpc_batch_payload_device_ptr(lpc_batch* b, int32_t* n_int64);

/* ---- EPS-native call: the batched mode as a solver uses it ---------------------------------------------------------
 * Embarrassingly parallel search hands the engine ONE root store, a list of decision variables and a set of subproblem
 * ids (bit j of an id keeps the lower or the upper half of decision variable j, as lpc_batch_init_split). Nothing but
 * that crosses the host link on the way in; the subproblem stores are generated on the chip, one per thread group, run
 * to their fixpoint (the loop of tests/pir_test.cpp:60-62 per store, over the shared table of pir.hpp:182-195), and on
 * the way out come the per-subproblem flags, the reduction record and ONLY the stores that did not fail, compacted -
 * is_extractable / extract (pir.hpp:873-898) for the whole batch. In LPC_MODE_AUTO (default) the propagators that
 * PIR::ask (pir.hpp:417-438) already entails on the root are dropped from the table first (what deinterpret's
 * remove_entailed does, pir.hpp:912-925): every subproblem is a tightening of the root, so they stay entailed and can
 * never change a store; and when the root is itself a common fixpoint of the table, the first sweep of a subproblem
 * evaluates only the propagators that mention a decision variable (nothing else can move: every other operand is as in
 * the root). LPC_MODE_SWEEP does neither (every sweep evaluates every propagator, the reference's work unit).
 * The model must fit the shared memory of an SM (<= 8191 variables, table + store slots <= 227 KB). */
typedef struct lpc_eps lpc_eps;
typedef struct lpc_eps_result {
  int64_t n_bot, n_solution, n_unknown;   /* as lpc_batch_result */
  int32_t best_bound;                     /* min over non-failed stores of lb(objective_var); INT32_MAX if none */
  int32_t max_sweeps_seen;
  int64_t sweeps_total;
  int64_t deductions;                     /* deduce() evaluations executed */
  int64_t n_survivors;                    /* non-failed stores (may exceed the survivor capacity: then only that many were kept) */
  int32_t n_live_records;                 /* propagators left in the table after dropping those entailed on the root */
  float device_ms;
  int32_t overflow_hazard;                /* as lpc_batch_result */
  int32_t n_first_sweep_records;          /* > 0: the root was a common fixpoint of the table, and the first sweep of every
                                             subproblem ran on these records only (those that mention a decision variable) */
} lpc_eps_result;

int lpc_eps_create(const lpc_table* t, int32_t max_subproblems, int32_t survivor_cap, lpc_eps** out);
int lpc_eps_destroy(lpc_eps* e);
int lpc_eps_set_rank(lpc_eps* e, int32_t rank, int32_t world);
void* lpc_eps_payload_device_ptr(lpc_eps* e, int32_t* n_int64);
/* Everything in one call, from and to HOST buffers: root_lbub (nvars pairs), decision_vars, ids (n, or NULL for
 * first_id + k) in; flags (n bytes: bit0 = bot, bit1 = all propagators entailed), up to max_survivors non-failed stores
 * (nvars pairs each) with their subproblem index k in [0, n) out (in no particular order; *n_written of them). Any output
 * pointer may be NULL. Synchronous. */
int lpc_eps_solve_host(lpc_eps* e, const int32_t* root_lbub, const int32_t* decision_vars, int32_t n_decisions,
                       const int64_t* ids, int64_t first_id, int32_t n, const lpc_fixpoint_opts* o, int32_t objective_var,
                       uint8_t* flags, int32_t* survivors_lbub, int32_t* survivor_index, int32_t max_survivors,
                       int32_t* n_written, lpc_eps_result* r);
/* The same in steps, for a problem that stays resident on the device between runs (bench.py's device-timed figure). What
 * is derived from the table, the root and the decision list alone - the packed table without the propagators entailed on
 * the root, the first-sweep table - is built by the first run after an upload and kept until the next upload. */
int lpc_eps_upload(lpc_eps* e, const int32_t* root_lbub, const int32_t* decision_vars, int32_t n_decisions,
                   const int64_t* ids, int64_t first_id, int32_t n);
int lpc_eps_run_async(lpc_eps* e, const lpc_fixpoint_opts* o, int32_t objective_var);
int lpc_eps_collect(lpc_eps* e, lpc_eps_result* r);
int lpc_eps_download(lpc_eps* e, uint8_t* flags, int32_t* survivors_lbub, int32_t* survivor_index, int32_t max_survivors,
                     int32_t* n_written);
/* Multi-GPU, one process per GPU of one node: the exchange of the reduction record fused into the batch kernel. Every rank
 * exports the CUDA IPC handle of its inbox (LPC_PEER_HANDLE_BYTES bytes), the handles are gathered by whatever plumbing the
 * processes share (torch.distributed in bench.py) and connected; from then on the kernel that finishes a rank's batch
 * writes the rank's record into every peer's inbox over NVLink, and every lpc_eps_run_async / lpc_eps_solve_host queues a
 * one-thread kernel behind it that waits for the peers' records and leaves the payload of lpc_eps_payload_device_ptr as an
 * all-reduce would - without a collective library call per step. Every rank must make the same sequence of calls (a call
 * only returns its payload once every peer has made it too), counted from lpc_eps_peer_connect on, and the ranks must be
 * synchronised once between the last connect and the first call (a barrier of the launching plumbing). A peer that never
 * arrives leaves payload[0] = -1 after a few seconds instead of hanging the GPU. lpc_eps_peer_connect fails with LPC_ERR_CUDA when two devices
 * have no peer access; the caller then keeps its all-reduce. */
#define LPC_PEER_HANDLE_BYTES 64
int lpc_eps_peer_export(lpc_eps* e, void* handle_out);
int lpc_eps_peer_connect(lpc_eps* e, int32_t rank, int32_t world, const void* handles);
/* Unmap the peers' inboxes (before the ranks destroy their handles: synchronise the ranks between the two). */
int lpc_eps_peer_disconnect(lpc_eps* e);
/* Sweeps each subproblem took (n ints), informational. */
int lpc_eps_sweeps(lpc_eps* e, int32_t* out);

/* ---- in-kernel search over the batch (snapshot / restore + branching next to the fixpoint) ---------
 * Every store of the batch is the root of a depth-first search run by ONE thread block: propagate (the fixpoint above),
 * then branch on the first non-singleton variable of `branch_vars` (input order) by bisection, lower half first; the
 * right branch is a store snapshot (PIR::snapshot, pir.hpp:857-861) pushed on a per-block stack in device memory and
 * restored (pir.hpp:863-870) on backtrack. A leaf (all branching variables fixed) is a solution iff every propagator is
 * entailed (is_extractable, pir.hpp:873-884). The stores of the batch are the roots and are left untouched. Counts are
 * schedule independent. A subproblem that would exceed max_depth snapshots or max_nodes nodes is abandoned and counted
 * in n_incomplete. */
typedef struct lpc_search_opts {
  int64_t max_nodes;      /* per subproblem, 0 = unlimited */
  int32_t max_depth;      /* snapshots per block stack (default 64) */
  int32_t objective_var;  /* < 0: none; else best_bound = min over solutions of lb(objective_var) */
  uint64_t stream;
  int32_t change_driven;  /* 1: a node's fixpoint starts from the propagators of the branched variable only and
                             re-evaluates 32-record groups as their variables change; 0: dense sweeps at every node;
                             -1 (default): change-driven for tables of 2,048 propagators or more, dense below */
  int32_t reserved;
} lpc_search_opts;

typedef struct lpc_search_result {
  int64_t n_solutions;
  int64_t n_nodes;          /* fixpoints computed */
  int64_t n_fails;          /* nodes whose fixpoint is bot */
  int64_t n_unknown_leaves; /* leaves with a propagator that is not entailed (branch_vars does not cover the model) */
  int64_t n_incomplete;     /* subproblems abandoned at a limit */
  int64_t sweeps_total;
  int64_t deductions;
  int32_t best_bound;       /* INT32_MAX if there is no solution or no objective */
  int32_t max_depth_seen;
  float device_ms;
  int32_t reserved;
} lpc_search_result;

void lpc_search_default_opts(lpc_search_opts* o);
/* per_store (may be NULL): n_stores x 6 int64 {solutions, nodes, fails, best, incomplete, unknown_leaves}. Synchronous. */
int lpc_batch_search(lpc_batch* b, const int32_t* branch_vars, int32_t n_branch, const lpc_search_opts* o,
                     lpc_search_result* r, int64_t* per_store);

#ifdef __cplusplus
}
#endif
#endif /* LPC_H */
