// grid_barrier.cuh — a grid-wide barrier that carries the sweep's votes, for the persistent fixpoint kernels.
//
// A sweep of the fixpoint ends with "did any block tighten a bound / see an empty variable" and a barrier. Doing that
// with a flag word plus cooperative_groups' grid.sync() costs four dependent L2 round trips per sweep (flag atomic,
// barrier atomic, barrier poll, flag read); small networks then spend most of a sweep there. Here the arrival IS the
// vote: one 64-bit atomicAdd per block adds 1 to the arrival count (bits 0-15), the block's `changed` count to bits
// 16-43 (1 per block for a plain vote; the number of record groups it changed for the change-driven kernel) and, if the
// block saw an empty variable, 1 to the `bot` count (bits 44-63); the value polled for the barrier already holds the
// grid's verdict. Three words rotate: while barrier k is in use, block 0 clears the word of barrier k + 1, which nobody
// can touch before block 0 itself has arrived at barrier k. Requires all blocks co-resident (cooperative launch).
#pragma once

namespace lpc {

struct GridVote { bool changed, bot; unsigned long long n_changed; };

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// words: 3 x u64 in global memory, zero at launch. `sweep` counts barriers from 0. Every thread of every block calls it.
// `block_changed` is block-uniform (already reduced over the block): how much this block adds to the changed count.
__device__ __forceinline__ GridVote grid_count_barrier(unsigned long long* words, int sweep, unsigned block_changed,
                                                       bool any_bot, unsigned long long* smem_slot) {
  const int slot = sweep % 3;
  __syncthreads();                                       // orders this block's joins before the arrival below
  if(threadIdx.x == 0) {
    if(blockIdx.x == 0) words[(sweep + 1) % 3] = 0ull;
    __threadfence();
    const unsigned long long add = 1ull | ((unsigned long long)block_changed << 16) | (any_bot ? 1ull << 44 : 0ull);
    unsigned long long v = atomicAdd(&words[slot], add) + add;
    const volatile unsigned long long* w = &words[slot];
    while((v & 0xffffull) != gridDim.x) v = *w;   // plain polling; the fence below is the acquire (and drops stale L1 lines)
    __threadfence();
    *smem_slot = v;
  }
  __syncthreads();
  const unsigned long long v = *smem_slot;
  GridVote r;
  r.n_changed = (v >> 16) & 0xfffffffull;
  r.changed = r.n_changed != 0;
  r.bot = (v >> 44) != 0;
  return r;
}

__device__ __forceinline__ GridVote grid_vote_barrier(unsigned long long* words, int sweep, bool changed, bool bot,
                                                      unsigned long long* smem_slot) {
  const bool any_changed = __syncthreads_or(changed);
  const bool any_bot = __syncthreads_or(bot);
  return grid_count_barrier(words, sweep, any_changed ? 1u : 0u, any_bot, smem_slot);
}

// SM-ordered block ranks. The kernels give block r the r-th contiguous fraction of the (y-sorted) table, so that the
// gathers of a block fall into one window of the store. Blocks that share an SM share its L1: with ranks handed out in
// SM order (all blocks of SM 0, then SM 1, ...) the co-resident blocks work on ADJACENT fractions and their windows
// overlap instead of tripling the SM's L1 footprint. Two steps around a grid barrier the caller already has:
//   slot = sm_rank_arrive(slots)        before the barrier: this block's arrival number on its SM
//   r = sm_rank_resolve(slots, slot)    after it: r.below + r.slot is a bijection onto [0, gridDim.x) whatever the
//                                       block-to-SM placement was; [r.below, r.below + r.here) is the SM's share.
// `slots`: LPC_SM_SLOTS ints in global memory, zero at launch. Both calls by every thread; the result is block-uniform.
#define LPC_SM_SLOTS 256
__device__ __forceinline__ unsigned my_smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
__device__ __forceinline__ int sm_rank_arrive(int* slots) {
  __shared__ int s_slot;
  if(threadIdx.x == 0) s_slot = atomicAdd(&slots[my_smid() & (LPC_SM_SLOTS - 1)], 1);
  __syncthreads();
  return s_slot;
}
struct SmRank { int below, here, slot; };   // blocks on lower-numbered SMs, blocks on this SM, this block's number among them
__device__ __forceinline__ SmRank sm_rank_resolve(const int* slots, int slot) {
  __shared__ int s_below, s_here;
  if(threadIdx.x < 32) {
    const int me = (int)(my_smid() & (LPC_SM_SLOTS - 1));
    int below = 0;
    for(int j = threadIdx.x; j < me; j += 32) below += reinterpret_cast<const volatile int*>(slots)[j];
    below = __reduce_add_sync(0xffffffffu, below);
    if(threadIdx.x == 0) { s_below = below; s_here = reinterpret_cast<const volatile int*>(slots)[me]; }
  }
  __syncthreads();
  SmRank r;
  r.below = s_below; r.here = s_here; r.slot = slot;
  return r;
}

} // namespace lpc
