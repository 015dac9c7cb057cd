// grid_barrier.cuh — a grid-wide barrier that carries the sweep's votes, for the persistent fixpoint kernels.
//
// A sweep of the fixpoint ends with "did any block tighten a bound / see an empty variable" and a barrier. Doing that
// with a flag word plus cooperative_groups' grid.sync() costs four dependent L2 round trips per sweep (flag atomic,
// barrier atomic, barrier poll, flag read); small networks then spend most of a sweep there. Here the arrival IS the
// vote: one 64-bit atomicAdd per block adds 1 to the arrival count (bits 0-19) and, if the block voted so, 1 to the
// `changed` count (bits 20-39) and to the `bot` count (bits 40-59); the value polled for the barrier already holds the
// grid's verdict. Three words rotate: while barrier k is in use, block 0 clears the word of barrier k + 1, which nobody
// can touch before block 0 itself has arrived at barrier k. Requires all blocks co-resident (cooperative launch).
#pragma once

namespace lpc {

struct GridVote { bool changed, bot; };

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// words: 3 x u64 in global memory, zero at launch. `sweep` counts barriers from 0. Every thread of every block calls it.
__device__ __forceinline__ GridVote grid_vote_barrier(unsigned long long* words, int sweep, bool changed, bool bot,
                                                      unsigned long long* smem_slot) {
  const int slot = sweep % 3;
  const bool any_changed = __syncthreads_or(changed);   // also orders this block's joins before the arrival below
  const bool any_bot = __syncthreads_or(bot);
  if(threadIdx.x == 0) {
    if(blockIdx.x == 0) words[(sweep + 1) % 3] = 0ull;
    __threadfence();
    const unsigned long long add = 1ull | (any_changed ? 1ull << 20 : 0ull) | (any_bot ? 1ull << 40 : 0ull);
    unsigned long long v = atomicAdd(&words[slot], add) + add;
    const volatile unsigned long long* w = &words[slot];
    while((v & 0xfffffull) != gridDim.x) v = *w;   // plain polling; the fence below is the acquire (and drops stale L1 lines)
    __threadfence();
    *smem_slot = v;
  }
  __syncthreads();
  const unsigned long long v = *smem_slot;
  GridVote r;
  r.changed = ((v >> 20) & 0xfffffull) != 0;
  r.bot = (v >> 40) != 0;
  return r;
}

} // namespace lpc
